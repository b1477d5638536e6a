// tcgen05 implicit-GEMM convolution for NHWC fp16 activations (sm_100a).
//
//   out[n, oh, ow, co] = act( scale[co] * sum_{kh,kw,ci} in[n, oh*s+kh-p, ow*s+kw-p, ci] * w[co, kh, kw, ci]
//                             + bias[co] (+ residual[n, oh, ow, co]) )
//
// GEMM view: M = output pixels (128 per tile = a TW x TH x TN box of the output), N = Cout, K = KH*KW*Cin.
// The A operand is never materialised: for every filter tap and 64-channel slice one 4-D TMA box
// {64 ch, TW, TH, TN} of the *input* tensor is loaded at the tap's offset; TMA's out-of-bounds zero fill
// implements the convolution padding, and stride-2 layers read from four "parity views" of the input
// (base pointer offset by (h&1, w&1), strides doubled), so the same box shape serves every layer.
// A box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle, which is exactly the K-major
// SWIZZLE_128B operand layout tcgen05.mma consumes.
//
// Replaces the cuDNN convolution + BatchNorm(eval) + ReLU/ReLU6 (+ residual add) call sites of the reference:
//   ACT/models/resnet.py:94-114 (Bottleneck.forward), ACT/models/mobilenet.py:32-68, ACT/models/ppo.py:33-39.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace af {

constexpr int kConvBlockM = 128;
constexpr int kConvBlockK = 64;          // fp16 elements per k-block = 128 B swizzle span
constexpr int kConvMaxBlockN = 256;
constexpr int kConvMaxStages = 8;
constexpr int kConvThreads = 576;        // warp0: TMA producer, warp1: MMA issuer, warps 2-17: four epilogue groups
constexpr int kConvSmemBudget = 227 * 1024;   // max dynamic shared memory per CTA on sm_100
constexpr int kConvStagingBytes = 128 * 128;  // one 128-row x 64-channel fp16 slice of the output tile

enum ConvAct : int { kActNone = 0, kActRelu = 1, kActRelu6 = 2 };

struct ConvKernelParams {
  // geometry of the output and of the tiling
  int N, Ho, Wo, Cout;
  int TW, TH, TN;                  // TW*TH*TN == 128
  int tiles_w, tiles_h, tiles_n;   // number of boxes along each output dim
  int n_blocks;                    // ceil(Cout / BN)
  int BN;                          // tile N (multiple of 16, <= 256)
  int KH, KW, stride, pad;
  int cblks;                       // ceil(Cin / 64)
  int stages;                      // smem pipeline depth
  // epilogue
  const float* scale;              // [n_blocks*BN]
  const float* bias;               // [n_blocks*BN]
  const __half* residual;          // NHWC, pixel stride res_stride, or nullptr
  long long res_stride;
  void* out;                       // NHWC, pixel stride out_stride (elements), fp16 or fp32
  long long out_stride;
  int out_f32;
  int act;
  int tma_store;                   // 1: fp16 output leaves through smem staging + TMA store (maps.out)
  int res_mma;                     // 1: residual tile is TMA-loaded (maps.res) and added by an identity-matrix MMA
  int wres;                        // 1: all weight k-blocks stay resident in shared memory (set by the launcher)
  int vhalo;                       // 1: one (TH+KH-1)-row input box per (kw, channel slice); vertical taps = row-shifted windows (maps.ah)
  int epi_bufs;                    // staging buffers per epilogue group (1)
  int epi_groups;                  // epilogue groups at work (4, or 2 for vertical-halo tiles; set by the launcher)
  int backoff_ns;                  // sleep between mbarrier probes of the single-thread roles (0 = spin on try_wait)
  int epi_backoff_ns;              // same for the epilogue warps waiting for an accumulator
  int debug_flags;                 // bring-up only (env AF_CONV_DEBUG): 1 = skip TMA stores
  // Temporal shift folded into the A-operand loads of a 1x1 conv (STH/ops/temporal_shift.py:29-46): the N images are
  // N / tsm_T clips of tsm_T frames; input channels [0, 16*tsm_f16) are read from frame t+1, the next 16*tsm_f16 from
  // frame t-1, the rest from frame t; frames outside the clip read as zeros (TMA out-of-bounds fill on the frame
  // axis of the 5-D map maps.a5).  0 = off.  TN divides tsm_T (a tile never straddles two clips).
  int tsm_T, tsm_f16;
  int k2_blocks;                   // > 0: a second 1x1 GEMM accumulates into the same tile before the epilogue -- the
                                   // projection shortcut of a bottleneck (maps.a2 = block input, strided; maps.b2 = its
                                   // weights): out = act(conv(in) + conv1x1_s(in2) + bias), k2_blocks = ceil(Cin2 / 64)
  int pool;                        // 1: stem mode -- MaxPool2d(3, 2, 1) fused into the epilogue, output through maps.pool
  int pair;                        // 1: clusters of two CTAs, cta_group::2 MMAs of M = 256 (conv_gemm.cu, PAIR); maps.b box = BN/2 rows
};

struct ConvTensorMaps {
  CUtensorMap a[4];   // parity views (index = (h&1)*2 + (w&1)); stride-1 layers use a[0] only
  CUtensorMap ah;     // vhalo: same tensor as a[0], box {64, TW, TH+KH-1, 1}
  CUtensorMap a2;     // second GEMM: input view {Cin2, Wo, Ho, N} (pixel (oh*s2, ow*s2) of the block input), box = a[0]'s
  CUtensorMap b2;     // second GEMM: packed weights [Cout_pad][K2_pad], box {64, BN (BN/2 per CTA of a pair)}
  CUtensorMap pool;   // pool mode: pooled output {Cout, Wo/2, Ho/2, N}, box {64, Wo/2, 1, 1}
  CUtensorMap a5;     // temporal-shift mode: the input as {C, W, H, T, clips}, box {64, TW, TH, TN, 1}
  CUtensorMap b;      // packed weights [Cout_pad][K_pad], K-major
  CUtensorMap out;    // output tensor {Cout, Wo, Ho, N}, box {64, TW, TH, TN} (only when tma_store)
  CUtensorMap res;    // residual tensor, same dims / box as `out` (only when res_mma)
};

cudaError_t launch_conv_gemm(const ConvTensorMaps& maps, const ConvKernelParams& p, int sm_count,
                             cudaStream_t stream);
size_t conv_gemm_smem_bytes(int BN, int res_mma, int wres_bytes, int stage_a_bytes, int epi_groups, int* stages_out,
                            int* epi_bufs_out);
bool conv_gemm_wres_ok(int n_blocks, int BN, int KH, int KW, int cblks, int pair = 0);
// Should this layer run as CTA pairs?  (env AF_CONV_PAIR: 0 = never, 1 = wherever legal, unset = heuristic)
bool conv_gemm_pair_ok(int N, int Ho, int Wo, int Cout, int BN, int K, int has_residual, int sm_count);

}  // namespace af
