// tcgen05 implicit-GEMM convolution kernel (see conv_gemm.cuh for the algorithm and reference call sites).
#include "conv_gemm.cuh"
#include "ptx.cuh"

#include <cstdlib>

namespace af {

using namespace ptx;

namespace {

struct __align__(8) ConvSmemCtrl {
  uint64_t full[kConvMaxStages];
  uint64_t empty[kConvMaxStages];
  uint64_t tmem_full[4];   // [stage] or, when single-slice tiles alternate between a stage's two groups, [stage + 2*sub]
  uint64_t tmem_empty[4];  // [stage]; single-slice tiles that alternate between four groups use four accumulators: [it & 3]
  uint64_t wfull;          // resident-weights mode: all weight k-blocks have landed
  uint32_t tmem_base;
  uint32_t pad;
};

constexpr int kStageABytes = kConvBlockM * kConvBlockK * 2;   // 16 KiB
constexpr int kCtrlBytes = 256;
constexpr int kIdentBytes = 64 * 128;                         // 64x64 fp16 identity, K-major, 128-B swizzled
constexpr int kEpilogueThreads = 128;                         // per epilogue group (4 warps = 4 TMEM lane quarters)
constexpr int kEpilogueGroups = 4;                            // groups g and g+2 drain TMEM accumulator stage g
constexpr int kAffineBytes = kEpilogueGroups * 2 * kConvMaxBlockN * 4;   // per group: scale[256] + bias[256] fp32
constexpr int kEpilogueBarrier = 1;                           // named barrier ids 1..4 (one per group)
static_assert(kConvThreads == 64 + kEpilogueGroups * kEpilogueThreads, "thread roles");

// single-thread roles (producer / MMA issuer) back off between probes so they do not steal issue slots from the
// epilogue warps that share their scheduler
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, int ns = 32) {
  while (!mbar_try_wait(bar, parity)) {
    if (ns > 0) __nanosleep(ns);
  }
}

// source frame offset of 16-channel step j under the temporal shift: +1 for the first fold, -1 for the second, else 0
__device__ __forceinline__ int tsm_source(int j, int f16) { return j < f16 ? 1 : (j < 2 * f16 ? -1 : 0); }

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == kActRelu) return fmaxf(x, 0.f);
  if (act == kActRelu6) return fminf(fmaxf(x, 0.f), 6.f);
  return x;
}

// Epilogue math for 32 accumulator columns of one row: y = act(acc * scale + bias) -> fp16, written as four 16-byte
// chunks into the 128-byte-swizzled staging row.  (A residual, when there is one, is already inside the accumulator:
// see the identity-matrix MMA in the issuer warp.)  ACT is compile-time so the hot loop has no per-element branches;
// ReLU / ReLU6 clamp on packed half2 after the conversion (exact: the clamp bounds 0 and 6 are representable and
// rounding is monotonic).
template <int ACT>
__device__ __forceinline__ void epilogue_half_slice(const uint32_t (&v)[32], const float* __restrict__ sc,
                                                    const float* __restrict__ bi, uint8_t* srow, int row,
                                                    int chunk0) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 sa = *reinterpret_cast<const float4*>(sc + g * 8), sb = *reinterpret_cast<const float4*>(sc + g * 8 + 4);
    const float4 ba = *reinterpret_cast<const float4*>(bi + g * 8), bb = *reinterpret_cast<const float4*>(bi + g * 8 + 4);
    float x[8];
    x[0] = fmaf(__uint_as_float(v[g * 8 + 0]), sa.x, ba.x);
    x[1] = fmaf(__uint_as_float(v[g * 8 + 1]), sa.y, ba.y);
    x[2] = fmaf(__uint_as_float(v[g * 8 + 2]), sa.z, ba.z);
    x[3] = fmaf(__uint_as_float(v[g * 8 + 3]), sa.w, ba.w);
    x[4] = fmaf(__uint_as_float(v[g * 8 + 4]), sb.x, bb.x);
    x[5] = fmaf(__uint_as_float(v[g * 8 + 5]), sb.y, bb.y);
    x[6] = fmaf(__uint_as_float(v[g * 8 + 6]), sb.z, bb.z);
    x[7] = fmaf(__uint_as_float(v[g * 8 + 7]), sb.w, bb.w);
    uint4 ov;
    __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __half2 h;
      if (ACT != kActNone) {
        // ReLU folded into the conversion (cvt.rn.relu.f16x2.f32; first source operand = upper half)
        uint32_t d;
        asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x[2 * j + 1]), "f"(x[2 * j]));
        h = *reinterpret_cast<__half2*>(&d);
      } else {
        h = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
      }
      if (ACT == kActRelu6) h = __hmin2(h, __float2half2_rn(6.f));
      oh2[j] = h;
    }
    *reinterpret_cast<uint4*>(srow + (((chunk0 + g) ^ (row & 7)) << 4)) = ov;     // 128-B swizzle, conflict-free
  }
}

__device__ __forceinline__ void epilogue_half_slice_act(int act, const uint32_t (&v)[32], const float* sc,
                                                        const float* bi, uint8_t* srow, int row, int chunk0) {
  if (act == kActRelu) epilogue_half_slice<kActRelu>(v, sc, bi, srow, row, chunk0);
  else if (act == kActRelu6) epilogue_half_slice<kActRelu6>(v, sc, bi, srow, row, chunk0);
  else epilogue_half_slice<kActNone>(v, sc, bi, srow, row, chunk0);
}

// Tile index -> (n-block, tile column, tile row, image group) as a mixed-radix counter advanced by a fixed step: the
// persistent loops move by gridDim.x (or 2 * gridDim.x) tiles per iteration, and decoding every tile with four
// integer divisions costs the single-thread roles ~1k cycles per tile -- more than a one-k-block tile's MMAs.
template <bool POOL = false>
struct TileCursor {
  int nb, tw, th, tn;
  int d_nb, d_tw, d_th, d_tn;
  __device__ __forceinline__ void init(int tile, int step, const ConvKernelParams& p) {
    nb = tile % p.n_blocks;
    int r = tile / p.n_blocks;
    tw = r % p.tiles_w;
    r /= p.tiles_w;
    th = r % p.tiles_h;
    tn = r / p.tiles_h;
    d_nb = step % p.n_blocks;
    r = step / p.n_blocks;
    d_tw = r % p.tiles_w;
    r /= p.tiles_w;
    d_th = r % p.tiles_h;
    d_tn = r / p.tiles_h;
  }
  __device__ __forceinline__ void advance(const ConvKernelParams& p) {
    if constexpr (POOL) {
      // pool mode: a wrap of the tile-row digit moves on by a whole grid of images
      if (++th == p.tiles_h) {
        th = 0;
        tn += d_tn;
      }
      return;
    }
    nb += d_nb;
    int c = nb >= p.n_blocks ? 1 : 0;
    nb -= c ? p.n_blocks : 0;
    tw += d_tw + c;
    c = tw >= p.tiles_w ? 1 : 0;
    tw -= c ? p.tiles_w : 0;
    th += d_th + c;
    c = th >= p.tiles_h ? 1 : 0;
    th -= c ? p.tiles_h : 0;
    tn += d_tn + c;
  }
  // pool mode (ConvKernelParams::pool): CTA `cta` of `ncta` walks the row-pair tiles of images cta, cta + ncta, ..
  // top to bottom, so that the epilogue can pool across consecutive tiles of an image
  __device__ __forceinline__ void init_pool(int cta, int ncta) {
    nb = tw = th = 0;
    tn = cta;
    d_nb = d_tw = d_th = 0;
    d_tn = ncta;
  }
};

// PAIR: the kernel runs as clusters of two CTAs (cta_group::2, see ptx.cuh): the pair shares every weight tile (each
// CTA keeps and loads only half of its BN rows) and one tcgen05.mma of M = 256 covers both CTAs' 128-pixel tiles --
// twice the work per issued instruction for the small-N layers whose bound is the MMA issue rate, half the weight
// traffic per CTA for the wide-N layers whose bound is the L2 -> SM operand rate.  The two tiles of a pair are
// image groups 2*tn and 2*tn + 1 of the same (n-block, tile row, tile column).
// TSM: temporal shift folded into the loads (ConvKernelParams::tsm_T); a template parameter so that the ordinary
// kernels' single-thread roles carry no trace of it.
// POOL: stem mode (ConvKernelParams::pool), only with VHALO; a template parameter for the same reason as TSM.
template <bool VHALO, bool PAIR, bool TSM, bool POOL = false>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvTensorMaps maps, const ConvKernelParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-B alignment is required by the 128-B swizzle; the dynamic smem window starts 1024-B aligned (no static
  // __shared__ in this kernel) -- trap rather than silently corrupt if that ever stops holding.
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();

  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs)
  const int bn_local = PAIR ? p.BN / 2 : p.BN;            // weight rows this CTA keeps in shared memory
  const int stage_b_bytes = bn_local * kConvBlockK * 2;
  // Resident weights (p.wres): a layer with a single n-block and a small filter keeps ALL its weight k-blocks in
  // shared memory for the life of the CTA; pipeline stages then carry the A operand only, which removes the
  // per-tile weight re-load (a third of the L2 -> SM traffic of the 64-channel layers) and deepens the A prefetch.
  // Vertical-halo mode (p.vhalo, resident weights only): a stage holds ONE box of TH + KH - 1 input rows per (kw,
  // 64-channel slice); the KH vertical taps are row-shifted windows of it -- a UMMA descriptor may start at any
  // 128-byte row of a swizzled buffer (tools/probe/umma_offset_probe.cu) -- so the tile's L2 -> SM operand traffic
  // drops from KH*KW to KW*(TH+KH-1)/TH boxes.
  const int stage_a_bytes = VHALO ? (p.TH + p.KH - 1) * p.TW * 128 : kStageABytes;
  const int stage_bytes = stage_a_bytes + (p.wres ? 0 : stage_b_bytes);   // multiples of 1024
  // [resident weights][stages][64x64 identity tile, only with res_mma][epilogue staging][scale/bias][barriers]
  uint8_t* wres = smem;
  smem += p.wres ? static_cast<size_t>(p.KH * p.KW * p.cblks) * stage_b_bytes : 0;
  uint8_t* ident = smem + static_cast<size_t>(p.stages) * stage_bytes;
  uint8_t* staging_base = ident + (p.res_mma ? kIdentBytes : 0);
  float* s_affine = reinterpret_cast<float*>(staging_base + p.epi_groups * kConvStagingBytes);
  ConvSmemCtrl* ctrl = reinterpret_cast<ConvSmemCtrl*>(reinterpret_cast<uint8_t*>(s_affine) + kAffineBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = p.tiles_w * p.tiles_h * (PAIR ? (p.tiles_n + 1) / 2 : p.tiles_n);
  const int total_tiles = m_tiles * p.n_blocks;
  const int tile0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);      // first tile of this CTA (pair)
  const int tstep = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  // tiles this CTA (pair) goes through; pool mode: whole images, tiles_h row-pair tiles each (evaluated inside the two
  // single-thread roles only: the epilogue warps are the register-critical path)
  auto count_iters = [&]() -> int {
    return POOL ? (static_cast<int>(blockIdx.x) < p.N
                       ? (p.N - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                             static_cast<int>(gridDim.x) * p.tiles_h
                       : 0)
                : (tile0 < total_tiles ? (total_tiles - tile0 + tstep - 1) / tstep : 0);
  };
  auto cursor_init = [&](TileCursor<POOL>& c) {
    if constexpr (POOL) c.init_pool(static_cast<int>(blockIdx.x), static_cast<int>(gridDim.x));
    else c.init(tile0, tstep, p);
  };
  const int num_kb = p.KH * p.KW * p.cblks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&ctrl->full[s], 1);
      mbar_init(&ctrl->empty[s], 1);
    }
    for (int a = 0; a < 4; ++a) mbar_init(&ctrl->tmem_full[a], 1);
    mbar_init(&ctrl->wfull, 1);
    for (int a = 0; a < 4; ++a) {
      // one arrive per warp of every epilogue group that reads the stage: both groups of the stage when a tile has
      // several 64-column slices (they split the slices), one group when it has a single slice (they alternate tiles)
      // (pair: the leader's barrier also takes the arrivals of the peer CTA's epilogue warps)
      mbar_init(&ctrl->tmem_empty[a],
                (((p.tma_store && p.BN <= 64) || (VHALO && p.epi_groups == 2)) ? 4 : 8) * (PAIR ? 2 : 1));
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
    if (p.tma_store) tma_prefetch_desc(&maps.out);
    if (p.res_mma) tma_prefetch_desc(&maps.res);
    if (VHALO) tma_prefetch_desc(&maps.ah);
    if (p.k2_blocks > 0) {
      tma_prefetch_desc(&maps.a2);
      tma_prefetch_desc(&maps.b2);
    }
    if (TSM) tma_prefetch_desc(&maps.a5);
    if (POOL) tma_prefetch_desc(&maps.pool);
    if (p.stride == 2) {
      tma_prefetch_desc(&maps.a[1]);
      tma_prefetch_desc(&maps.a[2]);
      tma_prefetch_desc(&maps.a[3]);
    }
  }
  if (warp == 1) {
    if constexpr (PAIR) {
      tmem_alloc_pair(&ctrl->tmem_base, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(&ctrl->tmem_base, 512);
      tmem_relinquish();
    }
  }
  if (p.res_mma && warp >= 2) {
    // B operand of the residual MMA: I[n][k] = (n == k), 64 rows of 128 B, 16-byte chunk c of row n at (c ^ (n & 7))
    // (pair: this CTA supplies rows [32 * rank, 32 * rank + 32) of the 64 x 64 identity, stored as rows 0..31)
    for (int i = threadIdx.x - 64; i < (PAIR ? 32 : 64) * 8; i += kEpilogueGroups * kEpilogueThreads) {
      const int row_l = i >> 3, c = i & 7;
      const int n = row_l + (PAIR ? 32 * static_cast<int>(rank) : 0);
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if ((n >> 3) == c) {
        const uint32_t one = (n & 1) ? 0x3C000000u : 0x00003C00u;   // fp16 1.0 in the high / low half
        const int wsel = (n & 7) >> 1;
        val.x = wsel == 0 ? one : 0u;
        val.y = wsel == 1 ? one : 0u;
        val.z = wsel == 2 ? one : 0u;
        val.w = wsel == 3 ? one : 0u;
      }
      *reinterpret_cast<uint4*>(ident + row_l * 128 + ((c ^ (row_l & 7)) << 4)) = val;
    }
    fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / TMA completion
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  // pair: completion bytes of both CTAs' loads are counted on the LEADER's full / wfull barriers
  const uint32_t full0_addr = PAIR ? mapa_u32(&ctrl->full[0], 0) : 0u;
  const uint32_t wfull_addr = PAIR ? mapa_u32(&ctrl->wfull, 0) : 0u;
  const uint32_t tmem_empty0_addr = PAIR ? mapa_u32(&ctrl->tmem_empty[0], 0) : 0u;
  (void)full0_addr;
  (void)wfull_addr;
  (void)tmem_empty0_addr;
  auto load2 = [&](void* dst, const CUtensorMap* m, int st, int c0, int c1) {
    if constexpr (PAIR) {
      tma_load_2d_elect_pair(dst, m, st < 0 ? wfull_addr : full0_addr + 8u * static_cast<uint32_t>(st), c0, c1);
    } else {
      tma_load_2d_elect(dst, m, st < 0 ? &ctrl->wfull : &ctrl->full[st], c0, c1);
    }
  };
  auto load4 = [&](void* dst, const CUtensorMap* m, int st, int c0, int c1, int c2, int c3) {
    if constexpr (PAIR) tma_load_4d_elect_pair(dst, m, full0_addr + 8u * static_cast<uint32_t>(st), c0, c1, c2, c3);
    else tma_load_4d_elect(dst, m, &ctrl->full[st], c0, c1, c2, c3);
  };
  auto load5 = [&](void* dst, const CUtensorMap* m, int st, int c0, int c1, int c2, int c3, int c4) {
    if constexpr (PAIR) tma_load_5d_elect_pair(dst, m, full0_addr + 8u * static_cast<uint32_t>(st), c0, c1, c2, c3, c4);
    else tma_load_5d_elect(dst, m, &ctrl->full[st], c0, c1, c2, c3, c4);
  };
  // the leader announces the bytes of BOTH CTAs; the peer's loads only complete on that barrier
  auto expect = [&](uint64_t* bar, uint32_t bytes) {
    if (!PAIR || rank == 0) mbar_arrive_expect_tx_elect(bar, PAIR ? 2u * bytes : bytes);
  };

  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap the tail of the previous
  // kernel in the stream; nothing below may read or write global memory before that kernel has completed.
  pdl_launch_dependents();
  if (warp == 0) {
    // ============================ TMA producer ============================
    {
      // converged warp: every lane runs the loop, one elected lane issues each TMA / expect_tx
      pdl_wait_prior_grid();
      int stage = 0;
      uint32_t phase = 0;
      const int bo = p.backoff_ns;
      const int brow = static_cast<int>(rank) * bn_local;   // pair: this CTA's half of every weight tile
      if (p.wres) {
        expect(&ctrl->wfull, static_cast<uint32_t>(num_kb * stage_b_bytes));
        for (int kb = 0; kb < num_kb; ++kb)
          load2(wres + static_cast<size_t>(kb) * stage_b_bytes, &maps.b, -1, kb * kConvBlockK, brow);
      }
      TileCursor<POOL> cur;
      cursor_init(cur);
      const int n_iter = count_iters();
      for (int pit = 0; pit < n_iter; ++pit, cur.advance(p)) {
        const int nb = cur.nb;
        const int tn_eff = PAIR ? cur.tn * 2 + static_cast<int>(rank) : cur.tn;
        const int ow0 = cur.tw * p.TW, oh0 = cur.th * p.TH, n0 = tn_eff * p.TN;
        int kb = 0;
        if constexpr (VHALO) {
          for (int kw = 0; kw < p.KW; ++kw) {
            for (int cb = 0; cb < p.cblks; ++cb) {
              mbar_wait_backoff(&ctrl->empty[stage], phase ^ 1, bo);
              uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
              expect(&ctrl->full[stage], static_cast<uint32_t>(stage_a_bytes));
              load4(sa, &maps.ah, stage, cb * kConvBlockK, ow0 + kw - p.pad, oh0 - p.pad, n0);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
        if constexpr (!VHALO && TSM) {
          // temporal shift folded into the loads: per 64-channel k-block one box per run of 16-channel steps that
          // share a source frame (t+1 / t-1 / t); the frame axis of maps.a5 zero-fills outside the clip
          const int tpc = p.tsm_T / p.TN;
          const int tb = tn_eff / tpc, t0 = (tn_eff - tb * tpc) * p.TN;
          for (int cb = 0; cb < p.cblks; ++cb) {
            int k = 0;
            while (k < 4) {
              const int src = tsm_source(4 * cb + k, p.tsm_f16);
              int ke = k + 1;
              while (ke < 4 && tsm_source(4 * cb + ke, p.tsm_f16) == src) ++ke;
              mbar_wait_backoff(&ctrl->empty[stage], phase ^ 1, bo);
              uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
              expect(&ctrl->full[stage], static_cast<uint32_t>(stage_bytes));
              load5(sa, &maps.a5, stage, cb * kConvBlockK, ow0, oh0, t0 + src, tb);
              if (!p.wres) load2(sa + kStageABytes, &maps.b, stage, cb * kConvBlockK, nb * p.BN + brow);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
              k = ke;
            }
          }
        }
        for (int kh = 0; kh < ((VHALO || TSM) ? 0 : p.KH); ++kh) {
          for (int kw = 0; kw < p.KW; ++kw) {
            int map_idx = 0, ch, cw;
            if (p.stride == 1) {
              ch = oh0 + kh - p.pad;
              cw = ow0 + kw - p.pad;
            } else {
              const int rh = kh - p.pad, rw = kw - p.pad;
              const int ph = rh & 1, pw = rw & 1;           // parity of the input row / column
              ch = oh0 + ((rh - ph) >> 1);                  // floor((kh-pad)/2) offset inside the parity view
              cw = ow0 + ((rw - pw) >> 1);
              map_idx = ph * 2 + pw;
            }
            for (int cb = 0; cb < p.cblks; ++cb, ++kb) {
              mbar_wait_backoff(&ctrl->empty[stage], phase ^ 1, bo);
              uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
              uint8_t* sb = sa + kStageABytes;
              expect(&ctrl->full[stage], static_cast<uint32_t>(stage_bytes));
              load4(sa, &maps.a[map_idx], stage, cb * kConvBlockK, cw, ch, n0);
              if (!p.wres) load2(sb, &maps.b, stage, kb * kConvBlockK, nb * p.BN + brow);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
        for (int kb2 = 0; kb2 < p.k2_blocks; ++kb2) {
          // second GEMM (projection shortcut): its k-blocks ride the same ring as (A box, weight box) stages
          mbar_wait_backoff(&ctrl->empty[stage], phase ^ 1, bo);
          uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
          expect(&ctrl->full[stage], static_cast<uint32_t>(stage_bytes));
          load4(sa, &maps.a2, stage, kb2 * kConvBlockK, ow0, oh0, n0);
          load2(sa + kStageABytes, &maps.b2, stage, kb2 * kConvBlockK, nb * p.BN + brow);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (p.res_mma) {
          // the residual tile rides the same pipeline as extra A boxes (64 output channels each)
          for (int j = 0; j * 64 < p.BN && nb * p.BN + j * 64 < p.Cout; ++j) {
            mbar_wait_backoff(&ctrl->empty[stage], phase ^ 1, bo);
            uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
            expect(&ctrl->full[stage], static_cast<uint32_t>(kStageABytes));
            load4(sa, &maps.res, stage, nb * p.BN + j * 64, ow0, oh0, n0);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (pair: the leader CTA only) ============================
    if (!PAIR || rank == 0) {
      // the whole warp runs the loop (converged); one elected lane issues each MMA / commit
      auto mma = [&](uint32_t d, uint32_t la, uint32_t lb, uint32_t id, uint32_t acc) {
        if constexpr (PAIR) umma_f16_ss_lo_elect_pair(d, la, lb, id, acc);
        else umma_f16_ss_lo_elect(d, la, lb, id, acc);
      };
      auto commit = [&](uint64_t* bar) {
        if constexpr (PAIR) umma_commit_elect_pair(bar);   // same barrier in both CTAs of the pair
        else umma_commit_elect(bar);
      };
      const uint32_t idesc = make_idesc_f16_f32(PAIR ? 2 * kConvBlockM : kConvBlockM, static_cast<uint32_t>(p.BN));
      const bool alternate_tiles = p.tma_store && p.BN <= 64 && !(VHALO && (p.epi_groups == 2 || POOL));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const int bo = p.backoff_ns;
      // descriptor low words: stage s of the A ring / k-block kb of the resident weights are fixed offsets apart
      const uint32_t a_lo0 = smem_desc_lo(smem_u32(smem)), a_step = static_cast<uint32_t>(stage_bytes) >> 4;
      const uint32_t w_lo0 = smem_desc_lo(smem_u32(wres)), b_step = static_cast<uint32_t>(stage_b_bytes) >> 4;
      if (p.wres) mbar_wait_backoff(&ctrl->wfull, 0);
      const int n_iter = count_iters();
      int mma_nb = tile0 % p.n_blocks;
      const int mma_dnb = tstep % p.n_blocks;
      for (; it < n_iter;
           ++it, mma_nb = mma_nb + mma_dnb >= p.n_blocks ? mma_nb + mma_dnb - p.n_blocks : mma_nb + mma_dnb) {
        // accumulator of this tile: two 256-column stages, or -- single-slice tiles handled by four alternating
        // epilogue groups -- four 128-column ones, so that four tiles are in flight between the MMA and the epilogue
        const int as = it & 1;
        const int ai = alternate_tiles ? (it & 3) : as;
        const uint32_t aphase = alternate_tiles ? (it >> 2) & 1 : (it >> 1) & 1;
        mbar_wait_backoff(&ctrl->tmem_empty[ai], aphase ^ 1, bo);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(alternate_tiles ? ai * 128 : as * kConvMaxBlockN);
        if constexpr (VHALO) {
          for (int kw = 0; kw < p.KW; ++kw) {
            for (int cb = 0; cb < p.cblks; ++cb) {
              mbar_wait_backoff(&ctrl->full[stage], phase, bo);
              tc_fence_after();
              const uint32_t la0 = a_lo0 + static_cast<uint32_t>(stage) * a_step;
              uint32_t lb = w_lo0 + static_cast<uint32_t>(kw * p.cblks + cb) * b_step;
              for (int kh = 0; kh < p.KH; ++kh, lb += b_step * static_cast<uint32_t>(p.KW * p.cblks)) {
                const uint32_t la = la0 + static_cast<uint32_t>(kh * p.TW) * 8u;   // kh rows down: kh * TW * 128 B
#pragma unroll
                for (int k = 0; k < kConvBlockK / 16; ++k)
                  mma(tmem_d, la + static_cast<uint32_t>(k * 2), lb + static_cast<uint32_t>(k * 2), idesc,
                                 (kw | cb | kh | k) != 0 ? 1u : 0u);
              }
              commit(&ctrl->empty[stage]);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
        if constexpr (!VHALO && TSM) {
          for (int cb = 0; cb < p.cblks; ++cb) {
            int k = 0;
            while (k < 4) {
              const int src = tsm_source(4 * cb + k, p.tsm_f16);
              int ke = k + 1;
              while (ke < 4 && tsm_source(4 * cb + ke, p.tsm_f16) == src) ++ke;
              mbar_wait_backoff(&ctrl->full[stage], phase, bo);
              tc_fence_after();
              const uint32_t la = a_lo0 + static_cast<uint32_t>(stage) * a_step;
              const uint32_t lb = p.wres ? w_lo0 + static_cast<uint32_t>(cb) * b_step : la + (kStageABytes >> 4);
              for (int kk = k; kk < ke; ++kk)
                mma(tmem_d, la + static_cast<uint32_t>(kk * 2), lb + static_cast<uint32_t>(kk * 2), idesc,
                                     (cb | kk) != 0 ? 1u : 0u);
              commit(&ctrl->empty[stage]);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
              k = ke;
            }
          }
        }
        for (int kb = 0; kb < ((VHALO || TSM) ? 0 : num_kb); ++kb) {
          mbar_wait_backoff(&ctrl->full[stage], phase, bo);
          tc_fence_after();
          const uint32_t la = a_lo0 + static_cast<uint32_t>(stage) * a_step;
          const uint32_t lb = p.wres ? w_lo0 + static_cast<uint32_t>(kb) * b_step : la + (kStageABytes >> 4);
#pragma unroll
          for (int k = 0; k < kConvBlockK / 16; ++k) {
            // advance 16 fp16 = 32 B inside the 128-B swizzle span: +2 in the (addr >> 4) field
            mma(tmem_d, la + static_cast<uint32_t>(k * 2), lb + static_cast<uint32_t>(k * 2), idesc,
                           (kb | k) != 0 ? 1u : 0u);
          }
          commit(&ctrl->empty[stage]);   // frees the smem slot once these MMAs have read it
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        for (int kb2 = 0; kb2 < p.k2_blocks; ++kb2) {
          mbar_wait_backoff(&ctrl->full[stage], phase, bo);
          tc_fence_after();
          const uint32_t la = a_lo0 + static_cast<uint32_t>(stage) * a_step;
          const uint32_t lb = la + (kStageABytes >> 4);
#pragma unroll
          for (int k = 0; k < kConvBlockK / 16; ++k)
            mma(tmem_d, la + static_cast<uint32_t>(k * 2), lb + static_cast<uint32_t>(k * 2), idesc, 1u);
          commit(&ctrl->empty[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (p.res_mma) {
          // acc[:, j*64 .. j*64+nj) += R_j (128 x 64 fp16) * I (64 x nj): the residual add, exact in fp32
          const int nb = mma_nb;
          const uint32_t li = smem_desc_lo(smem_u32(ident));
          for (int j = 0; j * 64 < p.BN && nb * p.BN + j * 64 < p.Cout; ++j) {
            const int nj = p.BN - j * 64 < 64 ? p.BN - j * 64 : 64;
            const uint32_t idesc_r = make_idesc_f16_f32(PAIR ? 2 * kConvBlockM : kConvBlockM, static_cast<uint32_t>(nj));
            mbar_wait_backoff(&ctrl->full[stage], phase, bo);
            tc_fence_after();
            const uint32_t la = a_lo0 + static_cast<uint32_t>(stage) * a_step;
#pragma unroll
            for (int k = 0; k < kConvBlockK / 16; ++k)
              mma(tmem_d + static_cast<uint32_t>(j * 64), la + static_cast<uint32_t>(k * 2),
                                   li + static_cast<uint32_t>(k * 2), idesc_r, 1u);
            commit(&ctrl->empty[stage]);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        // accumulator complete -> epilogue.  With single-slice tiles the stage's two groups take alternate tiles and
        // each waits on its own barrier, so that every waiter sees every phase of the barrier it polls.
        commit(&ctrl->tmem_full[alternate_tiles ? as + 2 * ((it >> 1) & 1) : as]);
      }
    }
  } else {
    // ============================ epilogue (4 groups x 4 warps) ============================
    // Groups g and g+2 only ever handle the tiles whose accumulator lives in TMEM stage g & 1 (local tile index
    // it = g&1, (g&1)+2, ..).  Inside a tile the two groups take alternate 64-column slices; when the tile is a single
    // slice they take alternate tiles instead.  Every group has its own staging buffer, scale/bias copy and named
    // barrier, so 16 warps keep the SM's four schedulers busy while the next tile's MMAs run.
    const int group = (warp - 2) >> 2;
    const int as = group & 1;                 // TMEM accumulator stage this group drains
    const int sub = group >> 1;               // which of the stage's two groups
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;      // row of the 128-row tile == TMEM lane
    const int et = threadIdx.x - 64 - group * kEpilogueThreads;   // 0..127 inside the group
    const uint32_t bar_id = kEpilogueBarrier + group;
    uint8_t* staging = staging_base + group * kConvStagingBytes;
    uint8_t* srow = staging + row * 128;
    float* g_scale = s_affine + group * 2 * kConvMaxBlockN;
    float* g_bias = g_scale + kConvMaxBlockN;
    int loaded_nb = -1;                       // n-block whose scale / bias currently sit in smem
    const int tw = row % p.TW;
    const int th = (row / p.TW) % p.TH;
    const int tn = row / (p.TW * p.TH);
    const int nslices = (p.BN + 63) >> 6;
    // epi_groups == 2 (MMA-heavy vertical-halo tiles): only groups 0 and 1 work, each draining every tile of its TMEM
    // stage alone; the staging buffers of groups 2 and 3 are given to the load pipeline instead
    const bool solo = VHALO && p.epi_groups == 2;
    const bool alternate_tiles = p.tma_store && nslices == 1 && !solo;
    const int sl_first = (alternate_tiles || solo) ? 0 : sub, sl_step = (alternate_tiles || solo) ? 1 : 2;
    pdl_wait_prior_grid();
    if constexpr (POOL) {
      // ---- stem mode: conv + BN + ReLU + MaxPool2d(3, stride 2, padding 1) (ACT/models/resnet.py:138-142, 213-216).
      // A tile is two full output rows (TW == Wo, TH == 2) and the CTA walks an image top to bottom, so pooled row t =
      // max over conv rows 2t-1 (second row of the previous tile), 2t, 2t+1 and columns 2x-1 .. 2x+1; the conv output
      // itself never leaves the SM.  Values are post-ReLU (>= 0): the padding of the pool is "max with 0".
      // The epilogue groups form a pipeline over a ring of three conv staging buffers: group 0 drains the accumulator of
      // tile t into buffer t % 3, groups 1 and 2 (256 threads, one pooled 16-byte chunk each) pool tile t from buffers
      // t % 3 and (t-1) % 3 and store the pooled row.
      constexpr uint32_t kBarFull = 5, kBarFree = 8;      // named barriers 5..7 (staged), 8..10 (buffer may be rewritten)
      constexpr uint32_t kBarPool = 11;                    // the two pooling groups among themselves
      constexpr uint32_t kPoolThreads = 2 * kEpilogueThreads, kPipeThreads = 3 * kEpilogueThreads;
      uint8_t* pool_st = staging_base + 3 * kConvStagingBytes;
      if (group == 0) {
        for (int i = et; i < kConvMaxBlockN; i += kEpilogueThreads) {
          const bool in = i < p.BN;
          g_scale[i] = (in && p.scale != nullptr) ? __ldg(p.scale + i) : (in ? 1.f : 0.f);
          g_bias[i] = in ? __ldg(p.bias + i) : 0.f;
        }
        named_barrier_sync(bar_id, kEpilogueThreads);
        int it = 0, b3 = 0;
        for (int n = blockIdx.x; n < p.N; n += gridDim.x) {
          for (int t = 0; t < p.tiles_h; ++t, ++it) {
            const int as_t = it & 1;
            uint8_t* cur_st = staging_base + b3 * kConvStagingBytes;
            mbar_wait_backoff(&ctrl->tmem_full[as_t], (it >> 1) & 1, p.epi_backoff_ns);
            tc_fence_after();
            // buffer b3 was the "previous tile" operand of the pooling of tile it-2: wait until group 1 is done with it
            if (it >= 3) asm volatile("bar.sync %0, %1;" ::"r"(kBarFree + static_cast<uint32_t>(b3)), "r"(kPipeThreads) : "memory");
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(as_t * kConvMaxBlockN);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              if (hf * 32 >= p.BN) break;
              uint32_t v[32];
              __syncwarp();
              tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(hf * 32), v);
              tmem_ld_wait();
              if ((hf + 1) * 32 >= p.BN) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctrl->tmem_empty[as_t]);
              }
              epilogue_half_slice_act(p.act, v, g_scale + hf * 32, g_bias + hf * 32, cur_st + row * 128, row, hf * 4);
            }
            __threadfence_block();
            asm volatile("bar.arrive %0, %1;" ::"r"(kBarFull + static_cast<uint32_t>(b3)), "r"(kPipeThreads) : "memory");
            b3 = b3 == 2 ? 0 : b3 + 1;
          }
        }
      } else if (group <= 2) {
        const int half_w = p.TW >> 1;
        const int pt = (group - 1) * kEpilogueThreads + et;   // 0..255 among the pooling threads
        int it = 0, b3 = 0;
        for (int n = blockIdx.x; n < p.N; n += gridDim.x) {
          for (int t = 0; t < p.tiles_h; ++t, ++it) {
            const int bprev = b3 == 0 ? 2 : b3 - 1;
            const uint8_t* cur_st = staging_base + b3 * kConvStagingBytes;        // conv rows 2t, 2t+1
            const uint8_t* prev_st = staging_base + bprev * kConvStagingBytes;    // conv rows 2t-2, 2t-1
            asm volatile("bar.sync %0, %1;" ::"r"(kBarFull + static_cast<uint32_t>(b3)), "r"(kPipeThreads) : "memory");
            if (pt == 0) tma_store_wait_read0();      // the previous pooled row has left the pool staging buffer
            named_barrier_sync(kBarPool, kPoolThreads);
            for (int item = pt; item < half_w * 8; item += kPoolThreads) {
              const int px = item >> 3, c = item & 7;   // pooled pixel, 16-byte channel chunk
              __half2 m[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) m[j] = __float2half2_rn(0.f);
#pragma unroll
              for (int dy = -1; dy <= 1; ++dy) {
                if (dy < 0 && t == 0) continue;
                const uint8_t* src = dy < 0 ? prev_st : cur_st;
                const int rr = dy < 0 ? 1 : dy;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                  const int x = 2 * px + dx;
                  if (x < 0) continue;
                  const int srow = rr * p.TW + x;
                  const uint4 q = *reinterpret_cast<const uint4*>(src + srow * 128 + ((c ^ (srow & 7)) << 4));
                  const __half2* qh = reinterpret_cast<const __half2*>(&q);
#pragma unroll
                  for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], qh[j]);
                }
              }
              *reinterpret_cast<uint4*>(pool_st + px * 128 + ((c ^ (px & 7)) << 4)) = *reinterpret_cast<const uint4*>(m);
            }
            // the "previous tile" buffer is no longer needed by anyone: group 0 may refill it (tile it+2)
            if (it >= 1) asm volatile("bar.arrive %0, %1;" ::"r"(kBarFree + static_cast<uint32_t>(bprev)), "r"(kPipeThreads) : "memory");
            fence_proxy_async();
            named_barrier_sync(kBarPool, kPoolThreads);
            if (pt == 0) {
              tma_store_4d(&maps.pool, pool_st, 0, 0, t, n);
              tma_store_commit();
            }
            b3 = b3 == 2 ? 0 : b3 + 1;
          }
        }
        if (pt == 0) tma_store_wait_all();
      }
    } else {
    TileCursor<false> cur;
    {
      const long long first = tile0 + static_cast<long long>(as) * tstep;
      cur.init(first < total_tiles ? static_cast<int>(first) : 0, 2 * tstep, p);
    }
    auto release_acc = [&](int ai) {   // hand a TMEM accumulator back to the MMA warp (of the leader CTA)
      if constexpr (PAIR) mbar_arrive_cluster(tmem_empty0_addr + 8u * static_cast<uint32_t>(ai));
      else mbar_arrive(&ctrl->tmem_empty[ai]);
    };
    for (int it = as; !(solo && sub == 1) && tile0 + static_cast<long long>(it) * tstep < total_tiles;
         it += 2, cur.advance(p)) {
      if (alternate_tiles && ((it >> 1) & 1) != sub) continue;
      const int nb = cur.nb;
      const int ow0 = cur.tw * p.TW, oh0 = cur.th * p.TH, n0 = (PAIR ? cur.tn * 2 + static_cast<int>(rank) : cur.tn) * p.TN;
      const int ow = ow0 + tw, oh = oh0 + th, n = n0 + tn;
      const bool valid = (ow < p.Wo) && (oh < p.Ho) && (n < p.N);
      const long long pix = (static_cast<long long>(n) * p.Ho + oh) * p.Wo + ow;
      const uint32_t aphase = alternate_tiles ? (it >> 2) & 1 : (it >> 1) & 1;
      uint64_t* acc_full = &ctrl->tmem_full[alternate_tiles ? as + 2 * sub : as];
      const int co_base = nb * p.BN;
      const int ai = alternate_tiles ? as + 2 * sub : as;   // accumulator index (== it & 3 when tiles alternate)
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(alternate_tiles ? ai * 128 : as * kConvMaxBlockN);
      if (p.tma_store) {
        // ---- fp16 output: TMEM -> registers -> swizzled smem slice (128 rows x 64 ch) -> TMA store
        if (nb != loaded_nb) {
          // per-channel scale / bias of this n-block -> smem (zero past BN so stray columns stay finite).  Every
          // thread of the group is past the previous tile's last barrier, i.e. past its last read of these arrays.
          for (int i = et; i < kConvMaxBlockN; i += kEpilogueThreads) {
            const bool in = i < p.BN;
            g_scale[i] = (in && p.scale != nullptr) ? __ldg(p.scale + co_base + i) : (in ? 1.f : 0.f);
            g_bias[i] = in ? __ldg(p.bias + co_base + i) : 0.f;
          }
          loaded_nb = nb;
          named_barrier_sync(bar_id, kEpilogueThreads);
        }
        mbar_wait_backoff(acc_full, aphase, p.epi_backoff_ns);
        tc_fence_after();
#pragma unroll 1
        for (int sl = sl_first; sl < nslices; sl += sl_step) {
          const int c0 = sl * 64;
          // a slice with at most 32 accumulator columns (BN = 32, or the tail of BN = 96 / 160) is one half-slice
          const int nhf = (p.BN - c0 > 32) ? 2 : 1;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            if (hf >= nhf) break;
            uint32_t v[32];
            __syncwarp();
            tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c0 + hf * 32), v);
            tmem_ld_wait();
            if (hf == nhf - 1 && sl + sl_step >= nslices) {
              // this group's share of the accumulator is in registers: hand the TMEM stage back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) release_acc(ai);
            }
            if (hf == 0) {
              // the group's staging buffer may still be feeding its previous TMA store
              if (et == 0) tma_store_wait_read0();
              named_barrier_sync(bar_id, kEpilogueThreads);
            }
            epilogue_half_slice_act(p.act, v, g_scale + c0 + hf * 32, g_bias + c0 + hf * 32, srow, row, hf * 4);
          }
          fence_proxy_async();
          named_barrier_sync(bar_id, kEpilogueThreads);
          if (et == 0 && !(p.debug_flags & 1)) {
            tma_store_4d(&maps.out, staging, co_base + c0, ow0, oh0, n0);
            tma_store_commit();
          }
        }
      } else {
        // ---- fp32 (or odd-shaped) output: direct global stores, one row per thread
        mbar_wait_backoff(acc_full, aphase, p.epi_backoff_ns);
        tc_fence_after();
        for (int c0 = sub * 16; c0 < p.BN; c0 += 32) {   // the stage's two groups take alternate 16-column chunks
          uint32_t v[16];
          __syncwarp();
          tmem_ld_32x32b_x16(taddr + static_cast<uint32_t>(c0), v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int co = co_base + c0 + g * 8;
              if (co >= p.Cout) continue;
              float x[8];
              const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
              const float4 s0 = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + co)) : one4;
              const float4 s1 = p.scale ? __ldg(reinterpret_cast<const float4*>(p.scale + co + 4)) : one4;
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + co));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + co + 4));
              x[0] = fmaf(__uint_as_float(v[g * 8 + 0]), s0.x, b0.x);
              x[1] = fmaf(__uint_as_float(v[g * 8 + 1]), s0.y, b0.y);
              x[2] = fmaf(__uint_as_float(v[g * 8 + 2]), s0.z, b0.z);
              x[3] = fmaf(__uint_as_float(v[g * 8 + 3]), s0.w, b0.w);
              x[4] = fmaf(__uint_as_float(v[g * 8 + 4]), s1.x, b1.x);
              x[5] = fmaf(__uint_as_float(v[g * 8 + 5]), s1.y, b1.y);
              x[6] = fmaf(__uint_as_float(v[g * 8 + 6]), s1.z, b1.z);
              x[7] = fmaf(__uint_as_float(v[g * 8 + 7]), s1.w, b1.w);
              const bool full8 = (co + 8 <= p.Cout);
              if (p.residual != nullptr) {
                const __half* rp = p.residual + pix * p.res_stride + co;
                for (int j = 0; j < 8 && co + j < p.Cout; ++j) x[j] += __half2float(rp[j]);
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) x[j] = apply_act(x[j], p.act);
              if (p.out_f32) {
                float* op = reinterpret_cast<float*>(p.out) + pix * p.out_stride + co;
                if (full8) {
                  reinterpret_cast<float4*>(op)[0] = make_float4(x[0], x[1], x[2], x[3]);
                  reinterpret_cast<float4*>(op)[1] = make_float4(x[4], x[5], x[6], x[7]);
                } else {
                  for (int j = 0; j < 8 && co + j < p.Cout; ++j) op[j] = x[j];
                }
              } else {
                __half* op = reinterpret_cast<__half*>(p.out) + pix * p.out_stride + co;
                if (full8) {
                  uint4 o4;
                  __half2* oh2 = reinterpret_cast<__half2*>(&o4);
#pragma unroll
                  for (int j = 0; j < 4; ++j) oh2[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
                  *reinterpret_cast<uint4*>(op) = o4;
                } else {
                  for (int j = 0; j < 8 && co + j < p.Cout; ++j) op[j] = __float2half_rn(x[j]);
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(ai);
      }
    }
    if (p.tma_store && et == 0) tma_store_wait_all();   // smem must stay valid until the last store has read it
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // no CTA leaves (or frees TMEM) while its peer may still signal it / read its operands
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool conv_gemm_wres_ok(int n_blocks, int BN, int KH, int KW, int cblks, int pair) {
  static const bool no_wres = getenv("AF_NO_WRES") != nullptr;
  const long long wbytes = 1LL * KH * KW * cblks * (pair ? BN / 2 : BN) * kConvBlockK * 2;   // per CTA
  return !no_wres && n_blocks == 1 && wbytes <= 80 * 1024;
}

bool conv_gemm_pair_ok(int N, int Ho, int Wo, int Cout, int BN, int K, int has_residual, int sm_count) {
  static const char* env = getenv("AF_CONV_PAIR");
  static const int mode = env ? atoi(env) : -1;      // -1 = heuristic
  if (mode == 0) return false;
  // legality: an even split of the weight tile in 8-row groups; the identity tile of the residual MMA is split per
  // 64-column slice, so those layers need whole slices; at least two image groups to pair up
  if (BN % 16 != 0 || N < 2) return false;
  if (has_residual && BN % 64 != 0) return false;
  if (mode == 1) return true;
  // heuristic (measured per layer at 1024 patches, profiles/r2_pair_ab.txt): the pair pays for MMA-heavy tiles --
  // K >= 512 and a tile at least 128 columns wide (-2 % .. -8 %: the weight half-tiles halve each CTA's L2 -> SM
  // operand traffic); short-K, store-bound layers lose (the two CTAs' epilogues gate each other through the shared
  // accumulator hand-back: 32 -> 96 @112^2 145 -> 248 us) and a cta_group::2 MMA of N <= 64 costs more than two
  // cta_group::1 MMAs (3x3 64 -> 64 @32^2: 99 -> 110 us), so the issue-bound small-N layers stay single-CTA.
  const long long tiles = (1LL * N * Ho * Wo + 127) / 128 * ((Cout + BN - 1) / BN);
  return K >= 512 && BN >= 128 && tiles >= 4LL * sm_count;
}

size_t conv_gemm_smem_bytes(int BN, int res_mma, int wres_bytes, int stage_a_bytes, int epi_groups, int* stages_out,
                            int* epi_bufs_out) {
  // BN = weight rows per CTA (half the tile's N for CTA pairs)
  const int stage_bytes = stage_a_bytes + (wres_bytes ? 0 : BN * kConvBlockK * 2);
  // one 16 KiB staging buffer per epilogue group (other groups compute while a group's TMA store drains)
  const int epi_bufs = 1;
  const int fixed = epi_groups * epi_bufs * kConvStagingBytes + kAffineBytes + kCtrlBytes +
                    (res_mma ? kIdentBytes : 0);
  int stages = (kConvSmemBudget - fixed - wres_bytes) / stage_bytes;
  if (stages > kConvMaxStages) stages = kConvMaxStages;
  if (stages < 2) stages = 2;
  if (stages_out) *stages_out = stages;
  if (epi_bufs_out) *epi_bufs_out = epi_bufs;
  return static_cast<size_t>(stages) * stage_bytes + fixed + wres_bytes;
}

cudaError_t launch_conv_gemm(const ConvTensorMaps& maps, const ConvKernelParams& p_in, int sm_count,
                             cudaStream_t stream) {
  static_assert(sizeof(ConvSmemCtrl) <= kCtrlBytes, "ctrl block too large");
  ConvKernelParams p = p_in;
  int stages = 0, epi_bufs = 1;
  // resident weights: single n-block and at most 80 KiB of weights (>= 4 A-only stages remain)
  const int bn_local = p.pair ? p.BN / 2 : p.BN;
  const int wbytes = p.KH * p.KW * p.cblks * bn_local * kConvBlockK * 2;
  p.wres = (p.k2_blocks == 0 && conv_gemm_wres_ok(p.n_blocks, p.BN, p.KH, p.KW, p.cblks, p.pair)) ? 1 : 0;
  if (!p.wres) p.vhalo = 0;
  const int stage_a_bytes = p.vhalo ? (p.TH + p.KH - 1) * p.TW * 128 : kStageABytes;
  static const bool four_groups = getenv("AF_VHALO_4GROUPS") != nullptr;
  p.epi_groups = (p.vhalo && p.tma_store && !four_groups) ? 2 : kEpilogueGroups;
  if (p.pool) {
    if (!p.vhalo || !p.wres || p.pair || p.tsm_T > 0 || p.BN > 64 || p.TH != 2 || p.TN != 1 || p.tiles_w != 1 ||
        !p.tma_store || p.res_mma || p.act == kActNone)
      return cudaErrorInvalidValue;
    p.epi_groups = 4;   // a ring of three conv staging buffers + the pooled-row staging buffer
  }
  const size_t smem =
      conv_gemm_smem_bytes(bn_local, p.res_mma, p.wres ? wbytes : 0, stage_a_bytes, p.epi_groups, &stages, &epi_bufs);
  if (p.vhalo && (stages < 2 || stage_a_bytes % 1024 != 0)) return cudaErrorInvalidValue;
  p.stages = stages;
  p.epi_bufs = epi_bufs;
  static const int dbg = getenv("AF_CONV_DEBUG") ? atoi(getenv("AF_CONV_DEBUG")) : 0;
  p.debug_flags = dbg;
  static const int backoff = getenv("AF_CONV_BACKOFF_NS") ? atoi(getenv("AF_CONV_BACKOFF_NS")) : 32;
  p.backoff_ns = backoff;
  static const int epi_backoff = getenv("AF_CONV_EPI_BACKOFF_NS") ? atoi(getenv("AF_CONV_EPI_BACKOFF_NS")) : 0;
  p.epi_backoff_ns = epi_backoff;
  // function attributes are per device: remember which devices have been configured
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev]) {
    using Kern = void (*)(const ConvTensorMaps, const ConvKernelParams);
    const Kern kerns[7] = {conv_gemm_kernel<false, false, false>, conv_gemm_kernel<true, false, false>,
                           conv_gemm_kernel<false, true, false>,  conv_gemm_kernel<true, true, false>,
                           conv_gemm_kernel<false, false, true>,  conv_gemm_kernel<false, true, true>,
                           conv_gemm_kernel<true, false, false, true>};
    for (Kern k : kerns) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kConvSmemBudget);
      if (e != cudaSuccess) return e;
    }
    if (dev >= 0 && dev < 64) attr_set[dev] = true;
  }
  int grid;
  if (p.pair) {
    const int pairs = p.tiles_w * p.tiles_h * ((p.tiles_n + 1) / 2) * p.n_blocks;
    const int max_pairs = sm_count / 2;
    grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
    if (grid < 2) grid = 2;
  } else {
    const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.n_blocks;
    grid = total_tiles < sm_count ? total_tiles : sm_count;
    if (grid < 1) grid = 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kConvThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  static const bool pdl = getenv("AF_NO_PDL") == nullptr;
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (p.pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;   // the two CTAs of a pair land on the two SMs of one TPC
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (p.pool) return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true, false, false, true>, maps, p);
  if (p.tsm_T > 0) {
    if (p.vhalo) return cudaErrorInvalidValue;
    if (p.pair) return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<false, true, true>, maps, p);
    return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<false, false, true>, maps, p);
  }
  if (p.pair) {
    if (p.vhalo) return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true, true, false>, maps, p);
    return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<false, true, false>, maps, p);
  }
  if (p.vhalo) return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<true, false, false>, maps, p);
  return cudaLaunchKernelEx(&cfg, conv_gemm_kernel<false, false, false>, maps, p);
}

}  // namespace af
