// Probe: a depthwise 3x3 whose input is read straight out of TMEM, one THREAD per channel (the accumulator of a
// transposed expand GEMM: lanes = expanded channels, columns = pixels), rolling the image rows through registers.
// Measures (a) the raw tcgen05.ld rate with 16-24 warps per SM and (b) cycles per row step of the stride-1 and
// stride-2 inner loops, including the fp16 stores of the result into a K-major swizzled operand tile.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../adafocus_b200/csrc/ptx.cuh"
using namespace af::ptx;

__device__ __forceinline__ float sat01(float x) {
  float y;
  asm("add.sat.f32 %0, %1, 0f00000000;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
  float y;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
__device__ __forceinline__ void st_h(uint8_t* p, float v) {
  *reinterpret_cast<__half*>(p) = __float2half_rn(v);
}

// stride 1: state rows a, b; loads rows c, d (16 columns each = 14 outputs + halo); two output rows
__device__ __forceinline__ void step_s1(uint32_t taddr, float (&a)[16], float (&b)[16], float (&c)[16], float (&d)[16],
                                        const float (&w)[9], uint8_t* out, int chan_off) {
  uint32_t vc[16], vd[16];
  tmem_ld_32x32b_x16(taddr, vc);
  tmem_ld_32x32b_x16(taddr + 64, vd);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    c[i] = sat01(__uint_as_float(vc[i]));
    d[i] = sat01(__uint_as_float(vd[i]));
  }
#pragma unroll
  for (int j = 0; j < 14; ++j) {
    float o1 = a[j] * w[0];
    o1 = fmaf(a[j + 1], w[1], o1);
    o1 = fmaf(a[j + 2], w[2], o1);
    o1 = fmaf(b[j], w[3], o1);
    o1 = fmaf(b[j + 1], w[4], o1);
    o1 = fmaf(b[j + 2], w[5], o1);
    o1 = fmaf(c[j], w[6], o1);
    o1 = fmaf(c[j + 1], w[7], o1);
    o1 = fma_sat(c[j + 2], w[8], o1);
    float o2 = b[j] * w[0];
    o2 = fmaf(b[j + 1], w[1], o2);
    o2 = fmaf(b[j + 2], w[2], o2);
    o2 = fmaf(c[j], w[3], o2);
    o2 = fmaf(c[j + 1], w[4], o2);
    o2 = fmaf(c[j + 2], w[5], o2);
    o2 = fmaf(d[j], w[6], o2);
    o2 = fmaf(d[j + 1], w[7], o2);
    o2 = fma_sat(d[j + 2], w[8], o2);
    // K-major swizzled operand rows: pixel row = 128 B, 16-byte chunk index XOR (pixel & 7)
    st_h(out + j * 128 + (chan_off ^ ((j & 7) << 4)), o1);
    st_h(out + (64 + j) * 128 + (chan_off ^ ((j & 7) << 4)), o2);
  }
}

// stride 2: state row a (= 2r-1); loads rows c (2r), d (2r+1), 16 columns each = 7 outputs (15 columns used)
__device__ __forceinline__ void step_s2(uint32_t taddr, float (&a)[16], float (&c)[16], float (&d)[16], const float (&w)[9],
                                        uint8_t* out, int chan_off) {
  uint32_t vc[16], vd[16];
  tmem_ld_32x32b_x16(taddr, vc);
  tmem_ld_32x32b_x16(taddr + 64, vd);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    c[i] = sat01(__uint_as_float(vc[i]));
    d[i] = sat01(__uint_as_float(vd[i]));
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    float o = a[2 * j] * w[0];
    o = fmaf(a[2 * j + 1], w[1], o);
    o = fmaf(a[2 * j + 2], w[2], o);
    o = fmaf(c[2 * j], w[3], o);
    o = fmaf(c[2 * j + 1], w[4], o);
    o = fmaf(c[2 * j + 2], w[5], o);
    o = fmaf(d[2 * j], w[6], o);
    o = fmaf(d[2 * j + 1], w[7], o);
    o = fma_sat(d[2 * j + 2], w[8], o);
    st_h(out + j * 128 + (chan_off ^ ((j & 7) << 4)), o);
  }
}

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) probe(long long* out, const float* wts, int iters, int nwarps) {
  extern __shared__ __align__(1024) uint8_t smem[];   // 2 x 16 KiB operand tiles
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  {   // fill the TMEM columns of this warp's lane quarter
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int nwq = blockDim.x / 128;
    for (int c = 0; c < 512; ++c)
      if (c % nwq == (warp >> 2))
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + lane_sel + c), "r"(0x3e000000u + c) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    const int quarter = warp & 3, strip = (warp >> 2) & 3;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    float w[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) w[i] = wts[i * 128 + quarter * 32 + lane];
    const int chan = quarter * 32 + lane;           // channel inside a 64-wide K chunk pair; two chunks of 64
    uint8_t* tile = smem + (chan >> 6) * 16384;
    const int chan_off = (chan & 63) * 2;
    if (MODE == 0) {
      uint32_t acc = 0;
      for (int it = 0; it < iters; ++it) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem + lane_sel + ((it * 32 + strip * 8) & 255), v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
      }
      if (acc == 0x12345u) out[1] = acc;
    } else if (MODE == 1) {
      float a[16], b[16], c[16], d[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = b[i] = 0.1f;
      for (int it = 0; it < iters; it += 2) {
        const uint32_t ta = tmem + lane_sel + (((it >> 1) & 1) * 256) + strip * 14;
        step_s1(ta, a, b, c, d, w, tile + strip * 16 * 128, chan_off);
        step_s1(ta + 128, c, d, a, b, w, tile + strip * 16 * 128, chan_off);
      }
      if (a[3] + b[5] == 123.f) out[1] = 1;
    } else if (MODE == 2) {
      float a[16], c[16], d[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = 0.1f;
      for (int it = 0; it < iters; it += 2) {
        const uint32_t ta = tmem + lane_sel + (((it >> 1) & 1) * 256) + strip * 14;
        step_s2(ta, a, c, d, w, tile + strip * 16 * 128, chan_off);
        step_s2(ta + 128, d, c, a, w, tile + strip * 16 * 128, chan_off);
      }
      if (a[3] == 123.f) out[1] = 1;
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (t == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  float* w;
  cudaMalloc(&d, 16);
  cudaMalloc(&w, 9 * 128 * 4);
  std::vector<float> hw(9 * 128, 0.05f);
  cudaMemcpy(w, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 2 * 16384 + 128 * 128;
  const int iters = 4000;
  auto run = [&](auto kern, int mode, int nw, int threads) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; ++rep) kern<<<148, threads, smem>>>(d, w, iters, nw);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d nw %d: %s\n", mode, nw, cudaGetErrorString(e)); exit(1); }
    long long h;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double clk = double(h) / iters;
    if (mode == 0) printf("nw %2d  tcgen05.ld x32             %8.1f clk/iter  -> %7.1f B/clk/SM\n", nw, clk, nw * 4096.0 / clk);
    if (mode == 1) printf("nw %2d  stride-1 step (28 out/thr) %8.1f clk/step  -> %6.2f outputs/clk/SM\n", nw, clk, nw * 32 * 28.0 / clk);
    if (mode == 2) printf("nw %2d  stride-2 step ( 7 out/thr) %8.1f clk/step  -> %6.2f inputs/clk/SM\n", nw, clk, nw * 32 * 28.0 / clk);
  };
  for (int nw : {4, 8, 12, 16}) { run(probe<0, 512>, 0, nw, 512); run(probe<1, 512>, 1, nw, 512); run(probe<2, 512>, 2, nw, 512); }
  run(probe<0, 640>, 0, 20, 640); run(probe<1, 640>, 1, 20, 640); run(probe<2, 640>, 2, 20, 640);
  run(probe<0, 768>, 0, 24, 768); run(probe<1, 768>, 1, 24, 768); run(probe<2, 768>, 2, 24, 768);
  return 0;
}
