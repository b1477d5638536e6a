// Probe: issue rate of small tcgen05.mma shapes whose A operand is a row-shifted window of one swizzled buffer
// (the tensor-core depthwise of mbconv_tc.cu).  One CTA, one issuing thread, clock64 around `iters` batches.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../adafocus_b200/csrc/ptx.cuh"
using namespace af::ptx;

__global__ void __launch_bounds__(640, 1) rate(long long* out, int mode, int iters, int spinners) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                  // 320 rows x 128 B
  uint8_t* sB = smem + 320 * 128;      // 9 tiles x 64 rows x 128 B
  __shared__ uint64_t bar, bar2, bar_never;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int done;
  const int t = threadIdx.x, warp = t >> 5;
  for (int i = t; i < (320 * 128 + 9 * 8192) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (t == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar_never, 1); done = 0; fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_base_s, 128); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (t == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    const uint32_t id16 = make_idesc_f16_f32(128, 16), id64 = make_idesc_f16_f32(128, 64), id32 = make_idesc_f16_f32(128, 32);
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {          // 9 taps x 4 groups, N=16 K=16, shifted A
        for (int tp = 0; tp < 9; ++tp) {
          const uint64_t da = make_smem_desc_sw128(a0 + ((tp / 3) * 16 + tp % 3) * 128);
          const uint64_t db = make_smem_desc_sw128(b0 + tp * 2048);
          for (int q = 0; q < 4; ++q) umma_f16_ss(tmem + q * 16, da + q * 2, db + q * 2, id16, tp != 0);
        }
      } else if (mode == 1) {   // 9 taps x 4 k-slices, N=64 K=16, shifted A
        for (int tp = 0; tp < 9; ++tp) {
          const uint64_t da = make_smem_desc_sw128(a0 + ((tp / 3) * 16 + tp % 3) * 128);
          const uint64_t db = make_smem_desc_sw128(b0 + tp * 8192);
          for (int q = 0; q < 4; ++q) umma_f16_ss(tmem, da + q * 2, db + q * 2, id64, (tp | q) != 0);
        }
      } else if (mode == 2) {   // 36 x N=64 K=16, aligned A (plain GEMM k-steps)
        const uint64_t da = make_smem_desc_sw128(a0);
        const uint64_t db = make_smem_desc_sw128(b0);
        for (int i = 0; i < 36; ++i) umma_f16_ss(tmem, da + (i & 3) * 2, db + (i & 3) * 2, id64, i != 0);
      } else if (mode == 3) {   // 36 x N=16 K=16, aligned A
        const uint64_t da = make_smem_desc_sw128(a0);
        const uint64_t db = make_smem_desc_sw128(b0);
        for (int i = 0; i < 36; ++i) umma_f16_ss(tmem + (i & 3) * 16, da + (i & 3) * 2, db + (i & 3) * 2, id16, i > 3);
      } else if (mode == 4) {   // 9 taps x 2 groups, N=32 K=16 (two diagonal blocks side by side would need K=32; rate only)
        for (int tp = 0; tp < 9; ++tp) {
          const uint64_t da = make_smem_desc_sw128(a0 + ((tp / 3) * 16 + tp % 3) * 128);
          const uint64_t db = make_smem_desc_sw128(b0 + tp * 4096);
          for (int q = 0; q < 2; ++q) umma_f16_ss(tmem + q * 32, da + q * 2, db + q * 2, id32, tp != 0);
        }
      } else if (mode == 6) {   // 36 x N=64 K=16 aligned, tcgen05.commit after every 4 (as a k-block pipeline does)
        const uint64_t da = make_smem_desc_sw128(a0);
        const uint64_t db = make_smem_desc_sw128(b0);
        for (int i = 0; i < 36; ++i) {
          umma_f16_ss(tmem, da + (i & 3) * 2, db + (i & 3) * 2, id64, i != 0);
          if ((i & 3) == 3) umma_commit(&bar2);
        }
      } else if (mode == 7) {   // 36 x N=256 K=16 aligned
        const uint32_t id256 = make_idesc_f16_f32(128, 128);
        const uint64_t da = make_smem_desc_sw128(a0);
        const uint64_t db = make_smem_desc_sw128(b0);
        for (int i = 0; i < 36; ++i) umma_f16_ss(tmem, da + (i & 3) * 2, db + (i & 3) * 2, id256, i != 0);
      } else if (mode == 5) {   // 36 x N=8 K=16, aligned A
        const uint32_t id8 = make_idesc_f16_f32(128, 8);
        const uint64_t da = make_smem_desc_sw128(a0);
        const uint64_t db = make_smem_desc_sw128(b0);
        for (int i = 0; i < 36; ++i) umma_f16_ss(tmem + (i & 7) * 8, da + (i & 3) * 2, db + (i & 3) * 2, id8, i > 7);
      }
      umma_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
    }
    long long t1 = clock64();
    out[0] = t1 - t0;
    done = 1;
  } else if (t >= 128 && t < 128 + spinners * 32) {
    // idle warps polling an mbarrier that never completes, like epilogue warps waiting for an accumulator
    while (!done) {
      if (mbar_try_wait(&bar_never, 0)) break;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 320 * 128 + 9 * 8192;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"9x4 N16 K16 shifted", "9x4 N64 K16 shifted", "36 N64 K16 aligned", "36 N16 K16 aligned", "9x2 N32 K16 shifted", "36 N8 K16 aligned", "36 N64 commit/4", "36 N128 K16 aligned"};
  const int per[] = {36, 36, 36, 36, 18, 36, 36, 36};
  for (int spinners : {0, 16})
  for (int mode = 0; mode < 8; ++mode) {
    const int iters = 2000;
    if (mode == 0) printf("-- %d polling warps\n", spinners);
    rate<<<1, 640, smem>>>(d, mode, iters, spinners);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    long long h;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-24s %8.1f clk / batch  %6.1f clk / mma (incl. commit+wait per batch)\n", names[mode], double(h) / iters, double(h) / iters / per[mode]);
  }
  return 0;
}
