// Probe: can a K-major SWIZZLE_128B UMMA operand start at a 128-byte (one row) granularity inside a larger swizzled
// buffer?  A is 256 rows x 64 fp16 written with the address-based swizzle (chunk ^ (row & 7)); the MMA reads 128 rows
// starting at row `o`, with the descriptor's base-offset field either 0 or (o & 7).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../adafocus_b200/csrc/ptx.cuh"
using namespace af::ptx;

__global__ void __launch_bounds__(128, 1) probe(const __half* A, const __half* B, float* D, int o, int use_base, int n_mma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // 256 rows x 128 B
  uint8_t* sB = smem + 256 * 128;     // 64 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int i = t; i < 256 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
  }
  for (int i = t; i < 64 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
  }
  if (t == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (t == 0) {
    uint64_t da = make_smem_desc_sw128(smem_u32(sA) + o * 128);
    if (use_base) da |= static_cast<uint64_t>(o & 7) << 49;
    const uint64_t db = make_smem_desc_sw128(smem_u32(sB));
    const uint32_t idesc = make_idesc_f16_f32(128, n_mma);
    if (n_mma == 64) {
      for (int k = 0; k < 4; ++k) umma_f16_ss(tmem, da + k * 2, db + k * 2, idesc, k != 0);
    } else {
      // block-diagonal form: four N=16, K=16 MMAs, group g uses K slice g of A and rows 16g.. of B -> D columns 16g..
      for (int g = 0; g < 4; ++g) {
        const uint64_t dbg = make_smem_desc_sw128(smem_u32(sB) + g * 16 * 128);
        umma_f16_ss(tmem + g * 16, da + g * 2, dbg + g * 2, idesc, 0);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<__half> hA(256 * 64), hB(64 * 64);
  std::vector<float> fA(256 * 64), fB(64 * 64);
  for (int r = 0; r < 256; ++r) for (int k = 0; k < 64; ++k) { fA[r * 64 + k] = float((r * 7 + k * 3) % 17 - 8); hA[r * 64 + k] = __float2half(fA[r * 64 + k]); }
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) { fB[n * 64 + k] = float((n * 5 + k * 11) % 13 - 6); hB[n * 64 + k] = __float2half(fB[n * 64 + k]); }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 256 * 128 + 64 * 128;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int offs[] = {0, 1, 2, 3, 5, 7, 8, 9, 16, 18, 19, 37, 100};
  std::vector<float> hD(128 * 64);
  for (int n_mma : {64, 16}) for (int use_base = 0; use_base < 2; ++use_base) for (int o : offs) {
    cudaMemset(dD, 0, 128 * 64 * 4);
    probe<<<1, 128, smem>>>(dA, dB, dD, o, use_base, n_mma);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("n=%d base=%d o=%d: CUDA error %s\n", n_mma, use_base, o, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hD.data(), dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
      float ref = 0.f;
      if (n_mma == 64) { for (int k = 0; k < 64; ++k) ref += fA[(m + o) * 64 + k] * fB[n * 64 + k]; }
      else { const int g = n / 16; for (int k = g * 16; k < g * 16 + 16; ++k) ref += fA[(m + o) * 64 + k] * fB[n * 64 + k]; }
      if (ref != hD[m * 64 + n]) ++bad;
    }
    printf("n_mma=%d base_field=%d o=%d: %s (%d mismatches)\n", n_mma, use_base, o, bad ? "MISMATCH" : "ok", bad);
  }
  return 0;
}
