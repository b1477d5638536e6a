// Fused MobileNet-V2 inverted-residual block, third generation (sm_100a): ALL THREE convolutions of the block run on
// the tensor core (ACT/models/mobilenet.py:42-68, InvertedResidual.forward):
//
//   TMA: input tile + halo {64 ch, BW, BH} -> smem (128-B swizzle = A operand of the expand GEMM)
//   tcgen05.mma: D1[halo pixels, 64 expanded channels] = X * W1^T                 (TMEM, double-buffered per chunk)
//   expand-epilogue warps: D1 -> +bias, ReLU6, zero outside the image -> fp16 tile E in smem, one 128-byte swizzled
//       row per halo pixel in raster order (stride 2: four parity planes)
//   tcgen05.mma: D3[m, 16g..16g+16) += E[m + tap offset, 16g..16g+16) * diag(w_tap[16g..16g+16))   (9 taps x 4 groups of
//       M128 N16 K16).  The depthwise conv is a sum of nine row-shifted copies of E scaled per channel: a UMMA
//       descriptor may start at any 128-byte row of a swizzled buffer (the XOR uses absolute address bits, see
//       tools/probe/umma_offset_probe.cu), so the shift is just the descriptor's start address, and the per-channel
//       scale is a block-diagonal B operand whose 16x16 blocks are rebuilt per chunk from a compact fp16 table.
//   depthwise-epilogue warps: D3 -> +bias, ReLU6 -> fp16 A2 tile (rows m = oy * pitch + ox, junk columns included)
//   tcgen05.mma: D2[m, Cout] += A2 * W2[:, chunk]^T                               (TMEM, accumulated over the chunks)
//   depthwise-epilogue warps: D2 -> +bias (+ residual) -> fp16 -> swizzled staging (valid pixels only) -> TMA store
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "mbconv_fused.cuh"   // MbTensorMaps

namespace af {

constexpr int kMb3Threads = 608;   // warps 0-2: TMA producer / expand MMA / depthwise+project MMA; 3-18: epilogue warps

struct Mb3Params {
  int N, H, W, Cin, Cexp, Cout, S, Ho, Wo;
  int TW, TH;                    // output tile; (TH-1)*pitch + TW <= 128
  int BW, BH, n_rows, Mtiles;    // input halo box, its pixel count and the number of 128-row MMA tiles covering it
  int pitch;                     // row pitch (pixels) of E as the depthwise MMA sees it: BW (S=1), TW+1 (S=2 planes)
  int plane_rows;                // S=2: rows per parity plane of E (== 4 mod 8); S=1: unused
  int e_rows;                    // rows (128 B each) per E buffer
  int m_max;                     // rows of the depthwise / project accumulators that can hold a valid pixel
  int tiles_w, tiles_h;
  int XB, EB, AB, WB;            // buffers of: input window, E, A2, depthwise weight tiles (1 or 2 each)
  int D3B, D2B;                  // TMEM buffers of the depthwise / project accumulators
  int nA;                        // epilogue warps per TMEM lane quarter on the expand side (4 - nA on the depthwise side)
  int nc;                        // 64-channel chunks of the expanded tensor
  int k1steps;                   // ceil(Cin / 16)
  int cout_pad;                  // N of the project MMA (multiple of 16, <= 64)
  const float* bias1;            // [nc*64]
  const float* dw_w;             // [9][nc*64] fp32, BN scale folded in (converted to fp16 inside the kernel)
  const float* bias2;            // [nc*64]
  const float* bias3;            // [cout_pad]
  const __half* residual;
  long long res_stride;
  int off_w1, off_w2, off_a2, off_out, off_e, off_wb, off_c, off_ctrl, smem;
};

// Fills tiling / smem layout from N, H, W, Cin, Cexp, Cout, S; false if unsupported.
bool mbconv3_plan(Mb3Params* p);
cudaError_t launch_mbconv3(const MbTensorMaps& maps, const Mb3Params& p, int sm_count, cudaStream_t stream);

}  // namespace af
