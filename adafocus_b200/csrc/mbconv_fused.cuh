// Fused MobileNet-V2 inverted-residual block (sm_100a): 1x1 expand + BN + ReLU6 -> depthwise 3x3 (stride 1|2) + BN +
// ReLU6 -> 1x1 project + BN (+ residual) as ONE kernel (ACT/models/mobilenet.py:42-68, InvertedResidual.forward).
// The expanded tensor (6x the block's input, the dominant HBM traffic of the glance network) never leaves the SM:
//
//   TMA: input tile + halo {64 ch, BW, BH} -> smem (128-B swizzle = A operand of the expand GEMM)
//   tcgen05.mma: D1[halo pixels, 64 expanded channels] = X * W1^T            (TMEM, double-buffered per channel chunk)
//   compute warps: D1 -> +bias, ReLU6, zero outside the image -> fp16 tile E in smem
//   compute warps: depthwise 3x3 over E (dw_strip.cuh) -> +bias, ReLU6 -> fp16 A2 tile (128-B swizzled K-major operand)
//   tcgen05.mma: D2[128 output pixels, Cout] += A2 * W2[:, chunk]^T          (TMEM, accumulated over the chunks)
//   compute warps: D2 -> +bias (+ residual) -> fp16 -> swizzled staging -> TMA store
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace af {

constexpr int kMbThreads = 608;   // warps 0-2: TMA producer / expand MMA / project MMA, 3-10: epilogues, 11-18: depthwise

struct MbParams {
  int N, H, W, Cin, Cexp, Cout, S, Ho, Wo;
  int TW, TH, strips;            // output tile (TW % 8 == 0, TW * TH <= 128) and depthwise strips of 4 (2 at stride 2) rows
  int BW, BH, n_rows, Mtiles;    // input halo box, its pixel count and the number of 128-row MMA tiles covering it
  int tiles_w, tiles_h;
  int XB, EB, e_bytes;           // input window buffers / expanded-tile buffers (1 or 2 each), bytes per E buffer
  int nc;                        // 64-channel chunks of the expanded tensor
  int k1steps;                   // ceil(Cin / 16), or ceil((Cin + 2) / 16) with the bias columns
  int bias_col;                  // >= 0: input channel pair (bias_col, bias_col+1) carries the constant 1 that multiplies the
                                 // bias columns of w1 (af_mbconv_desc.bias1_in_w1); -1: bias added in the epilogue
  int cout_pad;                  // N of the project MMA (multiple of 16, <= 64)
  const float* bias1;            // [nc*64] expand bias (BN folded, zero padded)
  const float* dw_w;             // [9][nc*64] depthwise weights with the BN scale folded in, zero padded
  const float* bias2;            // [nc*64]
  const float* bias3;            // [cout_pad]
  const __half* residual;        // NHWC (N, Ho, Wo, Cout) with pixel stride res_stride, or nullptr
  long long res_stride;
  int off_w1, off_w2, off_a2, off_out, off_e, off_f32, off_ctrl, smem;
};

struct MbTensorMaps {
  CUtensorMap x;     // input {Cin, W, H, N}, box {64, BW, BH, 1}, 128-B swizzle
  CUtensorMap w1;    // packed expand weights [nc*64][64] (K-major, BN scale folded), box {64, 64}
  CUtensorMap w2;    // packed project weights [cout_pad][nc*64], box {64, cout_pad}
  CUtensorMap out;   // output {Cout, Wo, Ho, N}, box {64, TW, TH, 1}, 128-B swizzle
};

// Fills the tiling / shared-memory layout fields of p from N, H, W, Cin, Cexp, Cout, S; false if unsupported.
bool mbconv_plan(MbParams* p);
cudaError_t launch_mbconv_fused(const MbTensorMaps& maps, const MbParams& p, int sm_count, cudaStream_t stream);

}  // namespace af
