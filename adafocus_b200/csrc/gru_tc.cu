// Persistent tensor-core GRU recurrence (see gru_tc.cuh): all T steps of h_t = GRU(x_t, h_{t-1}) for up to 64
// sequences in ONE cooperative launch.
#include "gru_tc.cuh"

#include "ptx.cuh"

namespace af {

using namespace ptx;

namespace {

constexpr int kUnits = 8;                 // hidden units per CTA -> 24 gate rows (+ 8 zero rows) = N 32
constexpr int kWTileBytes = 32 * 128;     // one k-block of the CTA's weight slice: 32 rows x 64 fp16, 128-B swizzled
constexpr int kATileBytes = 64 * 128;     // one k-block of h: 64 sequences x 64 fp16
constexpr int kEpiWarp0 = 4;              // warps 4 and 5 own TMEM lanes 0..63 = the 64 sequences

struct __align__(8) GruCtrl {
  uint64_t full[kGruTcStages], empty[kGruTcStages];
  uint64_t wfull, acc_full, acc_empty;
  uint32_t tmem_base, pad;
};

__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <bool SPLIT>
__global__ void __launch_bounds__(kGruTcThreads, 1)
gru_tc_kernel(const __grid_constant__ GruTcMaps maps, const GruTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int nkb = p.H >> 6;                       // 64-column k-blocks of one weight part / one h part
  uint8_t* s_whi = smem;
  uint8_t* s_wlo = smem + static_cast<size_t>(nkb) * kWTileBytes;            // SPLIT only
  uint8_t* s_a = smem + static_cast<size_t>(SPLIT ? 2 : 1) * nkb * kWTileBytes;
  // (+ one tile of slack behind the ring: an M = 128 MMA reads 64 rows past its 64-row tile; those accumulator rows
  // are never looked at)
  GruCtrl* ctrl = reinterpret_cast<GruCtrl*>(s_a + (kGruTcStages + 1) * kATileBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u0 = blockIdx.x * kUnits;
  const unsigned int G = gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kGruTcStages; ++s) {
      mbar_init(&ctrl->full[s], 1);
      mbar_init(&ctrl->empty[s], 1);
    }
    mbar_init(&ctrl->wfull, 1);
    mbar_init(&ctrl->acc_full, 1);
    mbar_init(&ctrl->acc_empty, 2);
    fence_mbar_init();
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.h[0]);
    tma_prefetch_desc(&maps.h[1]);
  }
  if (warp == 1) {
    tmem_alloc(&ctrl->tmem_base, 32);
    tmem_relinquish();
  }
  // rows 24..31 of every weight tile (the fourth 8-row swizzle atom) are zero; the slack tile behind the ring too
  for (int i = threadIdx.x; i < (SPLIT ? 2 : 1) * nkb * 64; i += blockDim.x) {
    uint8_t* tile = smem + static_cast<size_t>(i >> 6) * kWTileBytes + 3 * 1024;
    reinterpret_cast<uint4*>(tile)[i & 63] = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int i = threadIdx.x; i < (kGruTcStages + 1) * kATileBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctrl->tmem_base;

  if (warp == 0) {
    // ============================ TMA producer ============================
    // weights once: per k-block three 8-row boxes (gates r, z, n of this CTA's units) -> atoms 0..2 of the tile
    const uint32_t wbytes = static_cast<uint32_t>((SPLIT ? 2 : 1) * nkb * 3 * 1024);
    mbar_arrive_expect_tx_elect(&ctrl->wfull, wbytes);
    for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
      const int col0 = part == 0 ? 0 : 2 * p.H;       // packed weights: [W_hi | W_hi | W_lo] along K (engine.pack_conv_split)
      uint8_t* base = part == 0 ? s_whi : s_wlo;
      for (int kb = 0; kb < nkb; ++kb)
        for (int g = 0; g < 3; ++g)
          tma_load_2d_elect(base + static_cast<size_t>(kb) * kWTileBytes + g * 1024, &maps.w, &ctrl->wfull,
                            col0 + kb * 64, g * p.H + u0);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < p.T; ++t) {
      // h_{t-1} of ALL units must have been published by every CTA (grid barrier: arrival count (t + 1) * G)
      const unsigned int target = static_cast<unsigned int>(t + 1) * G;
      while (ld_acquire_gpu(p.counter) < target) {
      }
      asm volatile("fence.proxy.async;" ::: "memory");   // other CTAs' generic-proxy stores -> this CTA's TMA reads
      const CUtensorMap* hm = &maps.h[t & 1];
      for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
        for (int kb = 0; kb < nkb; ++kb) {
          while (!mbar_try_wait(&ctrl->empty[stage], phase ^ 1)) {
          }
          mbar_arrive_expect_tx_elect(&ctrl->full[stage], kATileBytes);
          tma_load_2d_elect(s_a + stage * kATileBytes, hm, &ctrl->full[stage], part * p.H + kb * 64, 0);
          if (++stage == kGruTcStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    const uint32_t idesc = make_idesc_f16_f32(128, 32);
    while (!mbar_try_wait(&ctrl->wfull, 0)) __nanosleep(32);
    tc_fence_after();
    const uint32_t a_lo0 = smem_desc_lo(smem_u32(s_a));
    const uint32_t whi0 = smem_desc_lo(smem_u32(s_whi)), wlo0 = smem_desc_lo(smem_u32(s_wlo));
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < p.T; ++t) {
      while (!mbar_try_wait(&ctrl->acc_empty, (static_cast<uint32_t>(t) & 1u) ^ 1u)) {
      }
      tc_fence_after();
      uint32_t acc = 0;
      for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
        for (int kb = 0; kb < nkb; ++kb) {
          while (!mbar_try_wait(&ctrl->full[stage], phase)) {
          }
          tc_fence_after();
          const uint32_t la = a_lo0 + static_cast<uint32_t>(stage) * (kATileBytes >> 4);
          const uint32_t lb = whi0 + static_cast<uint32_t>(kb) * (kWTileBytes >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16_ss_lo_elect(tmem, la + static_cast<uint32_t>(k * 2), lb + static_cast<uint32_t>(k * 2), idesc, acc);
            acc = 1;
          }
          if (SPLIT && part == 0) {      // the same h_hi tile against W_lo
            const uint32_t lc = wlo0 + static_cast<uint32_t>(kb) * (kWTileBytes >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss_lo_elect(tmem, la + static_cast<uint32_t>(k * 2), lc + static_cast<uint32_t>(k * 2), idesc, 1u);
          }
          umma_commit_elect(&ctrl->empty[stage]);
          if (++stage == kGruTcStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      umma_commit_elect(&ctrl->acc_full);
    }
  } else if (warp >= kEpiWarp0) {
    // ============================ gate math: thread = sequence b, 8 hidden units ============================
    const int b = (warp - kEpiWarp0) * 32 + lane;
    const bool live = b < p.B;
    const int hparts = SPLIT ? 2 : 1;
    float h[kUnits], bh[3][kUnits];
#pragma unroll
    for (int j = 0; j < kUnits; ++j) {
      h[j] = (live && p.h0 != nullptr) ? p.h0[static_cast<long long>(b) * p.H + u0 + j] : 0.f;
#pragma unroll
      for (int g = 0; g < 3; ++g) bh[g][j] = __ldg(p.b_hh + g * p.H + u0 + j);
    }
    auto publish = [&](__half* hbuf) {
      // this CTA's slice of the operand rows h16[b] = [hi | lo] for the next step's loads
      if (live) {
        __align__(16) __half hi[kUnits], lo[kUnits];
#pragma unroll
        for (int j = 0; j < kUnits; ++j) {
          hi[j] = __float2half_rn(h[j]);
          lo[j] = __float2half_rn(h[j] - __half2float(hi[j]));
        }
        __half* row = hbuf + static_cast<long long>(b) * (hparts * p.H) + u0;
        *reinterpret_cast<uint4*>(row) = *reinterpret_cast<const uint4*>(hi);
        if (SPLIT) *reinterpret_cast<uint4*>(row + p.H) = *reinterpret_cast<const uint4*>(lo);
      }
      __threadfence();
      named_barrier_sync(1, 64);
      if (threadIdx.x == kEpiWarp0 * 32) {
        __threadfence();
        atomicAdd(p.counter, 1u);
      }
    };
    publish(p.hbuf);                                        // h_{-1} into buffer 0
    for (int t = 0; t < p.T; ++t) {
      // x-side pre-activations of this step: independent of the recurrence, fetched while the MMAs run
      float xr[kUnits], xz[kUnits], xn[kUnits];
      if (live) {
        const float* x = p.xg + (static_cast<long long>(b) * p.T + t) * 3 * p.H + u0;
        *reinterpret_cast<float4*>(xr) = __ldg(reinterpret_cast<const float4*>(x));
        *reinterpret_cast<float4*>(xr + 4) = __ldg(reinterpret_cast<const float4*>(x + 4));
        *reinterpret_cast<float4*>(xz) = __ldg(reinterpret_cast<const float4*>(x + p.H));
        *reinterpret_cast<float4*>(xz + 4) = __ldg(reinterpret_cast<const float4*>(x + p.H + 4));
        *reinterpret_cast<float4*>(xn) = __ldg(reinterpret_cast<const float4*>(x + 2 * p.H));
        *reinterpret_cast<float4*>(xn + 4) = __ldg(reinterpret_cast<const float4*>(x + 2 * p.H + 4));
#pragma unroll
        for (int j = 0; j < kUnits; ++j) {
          xr[j] += bh[0][j];
          xz[j] += bh[1][j];
        }
      }
      while (!mbar_try_wait(&ctrl->acc_full, static_cast<uint32_t>(t) & 1u)) {
      }
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16), v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->acc_empty);
      if (live) {
#pragma unroll
        for (int j = 0; j < kUnits; ++j) {
          const float r = sigmoid_(xr[j] + __uint_as_float(v[j]));
          const float z = sigmoid_(xz[j] + __uint_as_float(v[8 + j]));
          const float n = tanhf(xn[j] + r * (__uint_as_float(v[16 + j]) + bh[2][j]));
          h[j] = (1.f - z) * n + z * h[j];
        }
        // h_t for the consumers after the recurrence: fp16 rows (policy) or split rows [hi | lo | hi] (classifier)
        __align__(16) __half hi[kUnits], lo[kUnits];
#pragma unroll
        for (int j = 0; j < kUnits; ++j) {
          hi[j] = __float2half_rn(h[j]);
          lo[j] = __float2half_rn(h[j] - __half2float(hi[j]));
        }
        __half* hs = p.hseq + (static_cast<long long>(b) * p.T + t) * p.hseq_stride + u0;
        *reinterpret_cast<uint4*>(hs) = *reinterpret_cast<const uint4*>(hi);
        if (SPLIT) {
          *reinterpret_cast<uint4*>(hs + p.H) = *reinterpret_cast<const uint4*>(lo);
          *reinterpret_cast<uint4*>(hs + 2 * p.H) = *reinterpret_cast<const uint4*>(hi);
        }
        if (t == p.T - 1 && p.h_out != nullptr) {
          float* o = p.h_out + static_cast<long long>(b) * p.H + u0;
          *reinterpret_cast<float4*>(o) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(h[4], h[5], h[6], h[7]);
        }
      }
      if (t + 1 < p.T) publish(p.hbuf + static_cast<long long>((t + 1) & 1) * p.B * hparts * p.H);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

}  // namespace

size_t gru_tc_smem_bytes(int H, int split) {
  return static_cast<size_t>(split ? 2 : 1) * (H / 64) * kWTileBytes + (kGruTcStages + 1) * kATileBytes + 256;
}

bool gru_tc_supported(int B, int H, int sm_count, int split) {
  return B >= 1 && B <= 64 && H % 64 == 0 && H >= 64 && H / kUnits <= sm_count &&
         gru_tc_smem_bytes(H, split) <= 227 * 1024;
}

cudaError_t launch_gru_tc(const GruTcMaps& maps, const GruTcParams& p, int split, cudaStream_t s) {
  static_assert(sizeof(GruCtrl) <= 256, "ctrl block too large");
  using Kern = void (*)(const GruTcMaps, const GruTcParams);
  Kern kern = split ? gru_tc_kernel<true> : gru_tc_kernel<false>;
  const size_t smem = gru_tc_smem_bytes(p.H, split);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), s);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.H / kUnits);
  cfg.blockDim = dim3(kGruTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the steps are separated by a grid barrier
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, maps, p);
}

}  // namespace af
