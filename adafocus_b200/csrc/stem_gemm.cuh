// Fused crop + 3-channel stem convolution as a tcgen05 implicit GEMM (see stem_gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace af {

struct StemKernelParams {
  const float* frames;     // (N,3,H,W) fp32 NCHW
  const int32_t* yx;       // one (y,x) per yx_div frames, or nullptr (whole frame, P == H)
  int yx_div;
  int N, H, W, P;
  int KH, KW, stride, pad;
  int Ho, Wo;
  int TW, TH;              // output tile, TW*TH == 128 (one image per tile)
  int tiles_w, tiles_h;
  int KB;                  // k-blocks of 64 (K = KH*KW*3 rounded up)
  int BN;                  // == Cout (multiple of 16, <= 64)
  const float* scale;      // folded BN, or nullptr
  const float* bias;
  int act;
};

struct StemTensorMaps {
  CUtensorMap b;     // packed weights [Cout_pad][KB*64], box {64, BN}
  CUtensorMap out;   // output {Cout, Wo, Ho, N} fp16 NHWC, box {64, TW, TH, 1}
};

bool stem_gemm_supported(const StemKernelParams& p);
size_t stem_gemm_smem_bytes();
cudaError_t launch_stem_gemm(const StemTensorMaps& maps, const StemKernelParams& p, int sm_count, cudaStream_t stream);

}  // namespace af
