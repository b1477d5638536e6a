// TMA-staged depthwise 3x3 convolution + folded BatchNorm + activation, NHWC fp16 (see dwconv_tma.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace af {

struct DwTmaParams {
  const float* w9c;      // fp32 [9][C]
  const float* scale;    // fp32 [C]
  const float* bias;     // fp32 [C]
  int C, act;
  int CB, n_cb;          // channels per CTA (TMA box inner dim) and C / CB
  int TW, TH, RO, strips, NB;   // output tile: TW x TH (= strips * RO) pixels of NB images
  int tiles_w, tiles_h, tiles_n;
  int in_bytes, out_bytes;      // shared-memory bytes of one input window / one output tile
  int threads, smem;
};

// Chooses the tiling for a layer; false when the shape is not supported (the caller falls back to the direct kernel).
bool dwconv_tma_plan(int N, int H, int W, int C, int stride, DwTmaParams* out);

// in_map: {C, W, H, N} fp16, box {CB, TW*S + (S==1 ? 2 : 1), TH*S + (S==1 ? 2 : 1), NB}, no swizzle;
// out_map: {C, Wo, Ho, N} fp16, box {CB, TW, TH, NB}, no swizzle.
cudaError_t launch_dwconv3x3_tma(const CUtensorMap& in_map, const CUtensorMap& out_map, const DwTmaParams& p,
                                 int stride, int sm_count, cudaStream_t stream);

}  // namespace af
