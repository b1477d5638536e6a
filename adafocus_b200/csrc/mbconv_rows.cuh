// Row-streaming fused MobileNet-V2 inverted-residual block (sm_100a), the large-batch form of mbconv_fused.cuh
// (ACT/models/mobilenet.py:42-68, InvertedResidual.forward with expand_ratio != 1).
//
// The expand GEMM is issued TRANSPOSED: D1[expanded channel (TMEM lane), pixel (TMEM column)] = W1 * X^T, so that one
// THREAD owns one expanded channel and reads whole image rows of it straight out of TMEM (tcgen05.ld 32x32b).  The
// depthwise 3x3 is then thread-local: the rows roll through registers as the CTA marches down the frame, the nine
// weights and both biases are per-thread constants, nothing is staged through shared memory and no second warp group
// has to hand tiles over.  Its fp16 result is written as the K-major A operand of the project GEMM
// (D2[pixel, cout] += A2 * W2^T), whose epilogue adds bias (+ residual) and stores through TMA.
//
//   TMA        : G input rows (128 / RP, or 64 / RP when shared memory is short) x RP pixels x 64 channels per step
//                -> smem (B operand of the expand GEMM)
//   tcgen05.mma: E[slot] (128 lanes x 64 columns, fp32) = W1[chunk] * X[64 pixels]^T   ring of six slots in TMEM, one
//                MMA group per (64 / RP image rows, 128-lane chunk), two issuing warps
//   dw warps   : warp = (lane quarter, chunk, 14-output column strip): rows -> +bias, ReLU6 -> 3x3 -> +bias, ReLU6 -> A2
//   tcgen05.mma: D2 += A2[chunk] * W2[chunk]^T
//   epilogue   : D2 -> +bias (+ residual) -> fp16 -> staging -> TMA store (rows of the frame as they complete)
//
// ReLU6 is computed as 6 * sat(x / 6) (`add.sat.f32` / `fma.rn.sat.f32`, one instruction, the second one free): the
// host folds 1/6 into W1 / bias1 / bias2 and 6 into W2.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace af {

constexpr int kMrMaxChunks = 3;   // 128-lane chunks of the expanded tensor held in TMEM at once (3 x 128 columns)
constexpr int kMrMaxWQ = 5;       // depthwise warps per TMEM lane quarter
constexpr int kMrMaxBufs = 6;     // A2 operand buffers

// Channel -> (chunk, lane) placement of the expanded tensor and the depthwise work list.  Filled by mbrows_layout()
// from (Cexp, strips per row); the host packer reads it back through af_mbconv_rows_layout so that weights and kernel
// agree by construction.
struct MrLayout {
  int nchunks;
  int WQ;                                  // tasks (= depthwise warps) per lane quarter
  int16_t lane_ch[kMrMaxChunks][128];      // expanded channel computed by TMEM lane l of chunk c, or -1
  int16_t lane_kpos[kMrMaxChunks][128];    // column of that channel inside the chunk's A2 / W2 K range (0..127)
  int ksteps[kMrMaxChunks];                // 16-wide K steps of the project GEMM for chunk c
  int a2_bytes[kMrMaxChunks];              // 16384 or 32768
  int warps[kMrMaxChunks];                 // depthwise warps working on chunk c
  int8_t task_chunk[4][kMrMaxWQ];          // [quarter][task] -> chunk
  int8_t task_strip[4][kMrMaxWQ];          // [quarter][task] -> column strip
};

struct MrParams {
  int N, H, W, Cin, Cexp, Cout, S, Ho, Wo;
  int RP, G, SPR, segs;          // TMEM columns per row segment, input rows per step, strips per segment, segments per row
  int OW, OWseg, OR;             // outputs per strip (14 | 7), outputs per segment row, output rows per step
  int SPF, SPI;                  // steps per frame segment, steps per A2 item (2 when a step fills half of the 128 rows)
  int k1steps, cout_pad;
  int XS, NB;                    // input stages, A2 buffers (two per chunk)
  int x_stage, a2_sub;           // bytes per input stage (G * RP pixels), bytes per 64-channel A2 sub-tile (G * RP rows)
  int a2_off[kMrMaxBufs], a2_cnt[kMrMaxBufs];
  int w2_off[kMrMaxChunks];
  MrLayout lay;
  const float* dwp;              // [nchunks][11][128]: 9 taps, bias1 / 6, bias2 / 6 per TMEM lane
  const float* bias3;            // [cout_pad]
  const __half* residual;        // NHWC (N, H, W, Cout) with pixel stride res_stride, or nullptr (stride 1 only)
  long long res_stride;
  long long* prof;               // reserved (debug hook af_debug_mbconv_rows_prof), unused by the kernel
  int off_w1, off_w2, off_a2, off_out, off_ctrl, smem;
};

struct MrTensorMaps {
  CUtensorMap x;     // input {Cin, W, H, N}, box {64, RP, G, 1}, 128-B swizzle
  CUtensorMap w1;    // expand weights [nchunks*128][64] (row = TMEM lane, K-major), box {64, 128}
  CUtensorMap w2;    // project weights [cout_pad][nchunks*128] (column = chunk * 128 + kpos), box {64, cout_pad}
  CUtensorMap out;   // output {Cout, Wo, Ho, N}, box {64, OWseg, OR * SPI, 1}, 128-B swizzle
  // stride 1, first step of a frame: its first output row closes the PREVIOUS frame (box of one row), the other G-1
  // rows open the new frame (TMA stores do not take negative coordinates)
  CUtensorMap out_rest;   // box {64, OWseg, G - 1, 1}
  CUtensorMap out_one;    // box {64, OWseg, 1, 1}
};

bool mbrows_layout(int cexp, int spr, MrLayout* lay);
// Fills the geometry / shared-memory fields of p from N, H, W, Cin, Cexp, Cout, S; false if the shape is not handled.
bool mbrows_plan(MrParams* p);
cudaError_t launch_mbconv_rows(const MrTensorMaps& maps, const MrParams& p, int sm_count, cudaStream_t stream);

}  // namespace af
