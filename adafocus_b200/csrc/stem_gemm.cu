// Fused crop + stem convolution (sm_100a): get_patch (ACT/models/utils.py:37-51) + Conv2d(3, Cout, KxK, stride s,
// pad) + BatchNorm(eval) + ReLU (ACT/models/resnet.py:138-140) as ONE tcgen05 implicit-GEMM kernel.
//
// The im2col matrix (K = KH*KW*3 = 147 for the 7x7 ResNet stem, padded to 3 k-blocks of 64) is never written to HBM:
// four producer warps read the fp32 NCHW window of a 128-pixel output tile at the crop offset the policy head left
// in device memory, convert it to fp16 in shared memory (zero outside the P x P patch = the conv padding), and
// assemble the A operand tiles directly in the 128-byte-swizzled K-major layout that tcgen05.mma consumes
// (generic-proxy stores + fence.proxy.async + mbarrier).  The weights (Cout x 192 fp16) are TMA-loaded once per CTA
// and stay resident.  MMA issue, TMEM double buffering and the TMA-store epilogue are those of conv_gemm.cu.
#include <cstdlib>

#include "conv_gemm.cuh"
#include "ptx.cuh"
#include "stem_gemm.cuh"

namespace af {

using namespace ptx;

namespace {

constexpr int kStemThreads = 576;        // warp0 TMA(B), warp1 MMA, warps2-5 / 6-9 epilogue groups, warps10-17 producers
constexpr int kProducerThreads = 256;
constexpr int kProducerWarps = kProducerThreads / 32;
constexpr int kRowsPerThread = 128 * 8 / kProducerThreads;   // A-tile rows each producer thread assembles per k-block
constexpr int kProducerBarrier = 3;      // named barrier id (1, 2 belong to the epilogue groups)
constexpr int kAStages = 6;
constexpr int kAStageBytes = 128 * 128;  // 128 rows x 64 fp16
constexpr int kMaxKB = 4;                // K <= 256
constexpr int kWinElems = 6144;          // fp32 window of one tile (3 ch x rows x span), 24 KiB, double-buffered

struct __align__(8) StemCtrl {
  uint64_t a_full[kAStages];
  uint64_t a_empty[kAStages];
  uint64_t b_full;
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ float act_f(float x, int act) {
  if (act == kActRelu) return fmaxf(x, 0.f);
  if (act == kActRelu6) return fminf(fmaxf(x, 0.f), 6.f);
  return x;
}

__global__ void __launch_bounds__(kStemThreads, 1)
stem_gemm_kernel(const __grid_constant__ StemTensorMaps maps, const StemKernelParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int kb_count = p.KB;
  uint8_t* s_b = smem;                                              // [KB][BN rows x 128 B]
  uint8_t* s_a = s_b + kMaxKB * kConvMaxBlockN * 128 / 4;           // BN <= 64 -> 4 x 8 KiB reserved (32 KiB)
  uint8_t* s_stage = s_a + kAStages * kAStageBytes;                 // 2 groups x 16 KiB output staging
  float* s_win = reinterpret_cast<float*>(s_stage + 2 * kConvStagingBytes);   // [2][kWinElems] fp32 windows
  short* s_off = reinterpret_cast<short*>(s_win + 2 * kWinElems);   // [KB*64] window offsets per k, -1 = zero
  float* s_scale = reinterpret_cast<float*>(s_off + kMaxKB * 64);
  float* s_bias = s_scale + 64;
  StemCtrl* ctrl = reinterpret_cast<StemCtrl*>(s_bias + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = p.tiles_w * p.tiles_h * p.N;
  const int rows_win = (p.TH - 1) * p.stride + p.KH;
  const int span = (p.TW - 1) * p.stride + p.KW;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(&ctrl->a_full[s], 1);
      mbar_init(&ctrl->a_empty[s], 1);
    }
    mbar_init(&ctrl->b_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&ctrl->tmem_full[a], 1);
      mbar_init(&ctrl->tmem_empty[a], 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&maps.b);
    tma_prefetch_desc(&maps.out);
  }
  if (warp == 1) {
    tmem_alloc(&ctrl->tmem_base, 128);
    tmem_relinquish();
  }
  // per-k window offsets and the folded-BN affine (launch constants)
  const int kreal = p.KH * p.KW * 3;
  for (int k = threadIdx.x; k < kb_count * 64; k += blockDim.x) {
    int off = -1;
    if (k < kreal) {
      const int tap = k / 3, c = k - tap * 3;
      const int kh = tap / p.KW, kw = tap - kh * p.KW;
      off = (c * rows_win + kh) * span + kw;
    }
    s_off[k] = static_cast<short>(off);
  }
  if (threadIdx.x < 64) {
    s_scale[threadIdx.x] = threadIdx.x < p.BN ? (p.scale ? p.scale[threadIdx.x] : 1.f) : 0.f;
    s_bias[threadIdx.x] = threadIdx.x < p.BN ? p.bias[threadIdx.x] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  pdl_launch_dependents();

  if (warp == 0) {
    // ============================ weights: loaded once, resident ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(&ctrl->b_full, static_cast<uint32_t>(kb_count * p.BN * 128));
      for (int kb = 0; kb < kb_count; ++kb) tma_load_2d(s_b + kb * p.BN * 128, &maps.b, &ctrl->b_full, kb * 64, 0);
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16_f32(128, static_cast<uint32_t>(p.BN));
      mbar_wait(&ctrl->b_full, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        while (!mbar_try_wait(&ctrl->tmem_empty[as], aphase ^ 1)) __nanosleep(32);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * 64);
        for (int kb = 0; kb < kb_count; ++kb) {
          while (!mbar_try_wait(&ctrl->a_full[stage], phase)) __nanosleep(32);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(s_a + stage * kAStageBytes));
          const uint64_t db = make_smem_desc_sw128(smem_u32(s_b + kb * p.BN * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(tmem_d, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                        (kb | k) != 0 ? 1u : 0u);
          umma_commit(&ctrl->a_empty[stage]);
          if (++stage == kAStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&ctrl->tmem_full[as]);
      }
    }
  } else if (warp < 10) {
    // ============================ epilogue (2 groups x 4 warps, one 64-column slice per tile) ============================
    const int group = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int et = threadIdx.x - 64 - group * 128;
    uint8_t* group_staging = s_stage + group * kConvStagingBytes;   // one buffer per group
    for (int it = group; blockIdx.x + static_cast<long long>(it) * gridDim.x < m_tiles; it += 2) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int tw_i = tile % p.tiles_w;
      const int th_i = (tile / p.tiles_w) % p.tiles_h;
      const int n = tile / (p.tiles_w * p.tiles_h);
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(as * 64);
      mbar_wait(&ctrl->tmem_full[as], aphase);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32b_x32(taddr, v0);
      tmem_ld_32x32b_x32(taddr + 32u, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->tmem_empty[as]);
      // the previous store of this group must have finished reading the staging buffer
      if (et == 0) tma_store_wait_read0();
      named_barrier_sync(1 + group, 128);
      uint8_t* srow = group_staging + row * 128;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 ov;
        __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = g * 8 + 2 * j;
          const uint32_t r0 = (c < 32) ? v0[c] : v1[c - 32];
          const uint32_t r1 = (c + 1 < 32) ? v0[c + 1] : v1[c + 1 - 32];
          oh2[j] = __floats2half2_rn(act_f(fmaf(__uint_as_float(r0), s_scale[c], s_bias[c]), p.act),
                                     act_f(fmaf(__uint_as_float(r1), s_scale[c + 1], s_bias[c + 1]), p.act));
        }
        *reinterpret_cast<uint4*>(srow + ((g ^ (row & 7)) << 4)) = ov;
      }
      fence_proxy_async();
      named_barrier_sync(1 + group, 128);
      if (et == 0) {
        tma_store_4d(&maps.out, group_staging, 0, tw_i * p.TW, th_i * p.TH, n);
        tma_store_commit();
      }
    }
    if (et == 0) tma_store_wait_all();
  } else {
    // ============================ A producers (4 warps): window staging + swizzled im2col tiles ============================
    pdl_wait_prior_grid();                       // the crop coordinates come from the previous kernel
    const int pt = threadIdx.x - 320;            // 0..kProducerThreads-1
    const int pw = pt >> 5;                      // producer warp
    const int c8 = pt & 7;                       // 16-byte chunk (8 k values) this thread assembles
    const int r0 = pt >> 3;                      // first of its rows (r0, r0 + kProducerThreads/8, ...)
    // launch-invariant per-thread tables: window offset of each of its rows, and of each of its k values per k-block
    int rbase[kRowsPerThread];
    uint32_t sw_off[kRowsPerThread];
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      const int row = r0 + i * (kProducerThreads / 8);
      const int tw = row % p.TW, th = row / p.TW;
      rbase[i] = (th * p.stride) * span + tw * p.stride;
      sw_off[i] = static_cast<uint32_t>(row * 128 + ((c8 ^ (row & 7)) << 4));
    }
    short koff[kMaxKB][8];
#pragma unroll
    for (int kb = 0; kb < kMaxKB; ++kb)
#pragma unroll
      for (int j = 0; j < 8; ++j) koff[kb][j] = kb < kb_count ? s_off[kb * 64 + c8 * 8 + j] : static_cast<short>(-1);
    int stage = 0;
    uint32_t phase = 0;
    // Window staging is asynchronous (cp.async, zero-fill outside the patch = conv padding): the fp32 window of the
    // NEXT tile is in flight while the current tile's A operands are assembled.
    auto issue_window = [&](int tile, float* win) {
      const int tw_i = tile % p.tiles_w;
      const int th_i = (tile / p.tiles_w) % p.tiles_h;
      const int n = tile / (p.tiles_w * p.tiles_h);
      int y0 = 0, x0 = 0;
      if (p.yx != nullptr) {
        const int e = n / p.yx_div;
        y0 = max(0, min(p.yx[2 * e], p.H - p.P));
        x0 = max(0, min(p.yx[2 * e + 1], p.W - p.P));
      }
      const int iy0 = th_i * p.TH * p.stride - p.pad, ix0 = tw_i * p.TW * p.stride - p.pad;
      const float* base = p.frames + static_cast<long long>(n) * 3 * p.H * p.W;
      for (int cr = pw; cr < 3 * rows_win; cr += kProducerWarps) {
        const int c = cr / rows_win, r = cr - c * rows_win;
        const int iy = iy0 + r;
        const bool row_ok = (iy >= 0) && (iy < p.P);
        const float* rowp = base + (static_cast<long long>(c) * p.H + (y0 + (row_ok ? iy : 0))) * p.W + x0;
        const uint32_t dst = smem_u32(win + cr * span);
        for (int xx = lane; xx < span; xx += 32) {
          const int ix = ix0 + xx;
          const bool ok = row_ok && ix >= 0 && ix < p.P;
          const float* src = rowp + (ok ? ix : 0);
          const uint32_t nbytes = ok ? 4u : 0u;      // src-size 0 -> the 4 destination bytes are zero-filled
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + xx * 4), "l"(src), "r"(nbytes)
                       : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int buf = 0;
    if (blockIdx.x < m_tiles) issue_window(blockIdx.x, s_win);
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x) {
      const float* win = s_win + buf * kWinElems;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      // this tile's window has landed for every producer thread, and the previous tile's assembly (which read the
      // other buffer) is complete, so the other buffer can be refilled
      named_barrier_sync(kProducerBarrier, kProducerThreads);
      if (tile + static_cast<int>(gridDim.x) < m_tiles) issue_window(tile + gridDim.x, s_win + (buf ^ 1) * kWinElems);
#pragma unroll
      for (int kb = 0; kb < kMaxKB; ++kb) {
        if (kb >= kb_count) break;
        if (pt == 0) {
          while (!mbar_try_wait(&ctrl->a_empty[stage], phase ^ 1)) __nanosleep(32);
        }
        named_barrier_sync(kProducerBarrier, kProducerThreads);
        uint8_t* a_tile = s_a + stage * kAStageBytes;
#pragma unroll
        for (int i = 0; i < kRowsPerThread; ++i) {
          const float* wr = win + rbase[i];
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = koff[kb][j] >= 0 ? wr[koff[kb][j]] : 0.f;
          uint4 ov;
          __half2* o2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
          for (int j = 0; j < 4; ++j) o2[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
          *reinterpret_cast<uint4*>(a_tile + sw_off[i]) = ov;
        }
        fence_proxy_async();                     // generic-proxy smem writes -> visible to the tensor core (async proxy)
        named_barrier_sync(kProducerBarrier, kProducerThreads);
        if (pt == 0) mbar_arrive(&ctrl->a_full[stage]);
        if (++stage == kAStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      buf ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace

size_t stem_gemm_smem_bytes() {
  return static_cast<size_t>(kMaxKB) * kConvMaxBlockN * 128 / 4 + kAStages * kAStageBytes + 2 * kConvStagingBytes +
         2 * kWinElems * 4 + kMaxKB * 64 * 2 + 2 * 64 * 4 + 256;
}

bool stem_gemm_supported(const StemKernelParams& p) {
  const int rows_win = (p.TH - 1) * p.stride + p.KH;
  const int span = (p.TW - 1) * p.stride + p.KW;
  return p.BN >= 16 && p.BN <= 64 && p.BN % 16 == 0 && p.KB >= 1 && p.KB <= kMaxKB && p.TW * p.TH == 128 &&
         3 * rows_win * span <= kWinElems && p.stride >= 1 && p.stride <= 2;
}

cudaError_t launch_stem_gemm(const StemTensorMaps& maps, const StemKernelParams& p, int sm_count, cudaStream_t stream) {
  static bool attr_set[64] = {};   // function attributes are per device
  const size_t smem = stem_gemm_smem_bytes();
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(stem_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  const int m_tiles = p.tiles_w * p.tiles_h * p.N;
  int grid = m_tiles < sm_count ? m_tiles : sm_count;
  if (grid < 1) grid = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kStemThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = getenv("AF_NO_PDL") == nullptr;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, stem_gemm_kernel, maps, p);
}

}  // namespace af
