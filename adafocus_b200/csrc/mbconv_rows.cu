// Row-streaming fused inverted-residual block kernel (see mbconv_rows.cuh for the data flow and the reference call
// sites).
#include "mbconv_rows.cuh"

#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "ptx.cuh"

namespace af {

using namespace ptx;

namespace {

constexpr int kD2Col = 384;            // TMEM column of the project accumulators (2 x 64 columns)
constexpr int kFirstDwWarp = 8;        // warp 0: TMA producer, 1 and 3: expand MMAs, 2: project MMAs, 4-7: epilogue, 8..: depthwise
constexpr int kSlots = 6;              // E ring: six 64-column slots (64 / RP image rows of one chunk each); 384 columns
constexpr int kEpiBarrier = 1;

struct __align__(8) MrCtrl {
  uint64_t x_full[3], x_empty[3];
  uint64_t w_full;
  uint64_t e_full[kSlots], e_free[kSlots];   // slot s belongs to chunk s % nchunks
  uint64_t a2_full[kMrMaxBufs], a2_free[kMrMaxBufs];
  uint64_t d2_full[2], d2_free[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void wait_sleep(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(32);
}
__device__ __forceinline__ void wait_spin(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// x + b clamped to [0, 1]
__device__ __forceinline__ float add_sat(float x, float b) {
  float y;
  asm("add.sat.f32 %0, %1, %2;" : "=f"(y) : "f"(x), "f"(b));
  return y;
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
  float y;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
__device__ __forceinline__ void st_half(uint32_t saddr, float v) {
  const unsigned short h = __half_as_ushort(__float2half_rn(v));
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(saddr), "h"(h) : "memory");
}

// One row of 16 columns out of TMEM -> +bias, clamp to [0, 1]; edge columns and rows outside the frame -> 0 (the
// depthwise zero padding lives in the expanded domain).
__device__ __forceinline__ void load_row(uint32_t taddr, float b1, bool zero_first, bool zero_last, bool zero_row,
                                         float (&r)[16]) {
  uint32_t v[16];
  tmem_ld_32x32b_x16(taddr, v);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = add_sat(__uint_as_float(v[i]), b1);
  if (zero_first) r[0] = 0.f;
  if (zero_last) r[15] = 0.f;
  if (zero_row) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = 0.f;
  }
}

// stride 1: one output row of 14 columns from input rows (r0, r1, r2).  MODE 1: r2 opens a new frame, so the row
// closes the previous frame (r2 counts as zero); MODE 2: the row is the first of a new frame (r0 counts as zero).
// ob = address of the row's first pixel in the A2 tile + this lane's K-column offset; the 16-byte chunk index is XORed
// with (pixel & 7) (128-B swizzle; the strip starts at a multiple of 16 pixels).
template <int MODE>
__device__ __forceinline__ void dw_row_s1(const float (&r0)[16], const float (&r1)[16], const float (&r2)[16],
                                          const float (&w)[9], float b2, uint32_t ob) {
#pragma unroll
  for (int j = 0; j < 14; ++j) {
    float o = b2;
    if (MODE != 2) {
      o = fmaf(r0[j], w[0], o);
      o = fmaf(r0[j + 1], w[1], o);
      o = fmaf(r0[j + 2], w[2], o);
    }
    o = fmaf(r1[j], w[3], o);
    o = fmaf(r1[j + 1], w[4], o);
    if (MODE == 1) {
      o = fma_sat(r1[j + 2], w[5], o);
    } else {
      o = fmaf(r1[j + 2], w[5], o);
      o = fmaf(r2[j], w[6], o);
      o = fmaf(r2[j + 1], w[7], o);
      o = fma_sat(r2[j + 2], w[8], o);
    }
    st_half((ob ^ static_cast<uint32_t>((j & 7) << 4)) + j * 128, o);
  }
}

// stride 2: one output row (a, c, d) of 7 columns; a = input row 2r-1, c = 2r, d = 2r+1.  Two phases, so that the
// taps of rows a and c run while the GEMM of row d may still be in flight (the E ring holds three rows per chunk).
__device__ __forceinline__ void dw_row_s2_ac(const float (&a)[16], const float (&c)[16], const float (&w)[9], float b2,
                                             float (&o)[7]) {
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    float v = fmaf(a[2 * j], w[0], b2);
    v = fmaf(a[2 * j + 1], w[1], v);
    v = fmaf(a[2 * j + 2], w[2], v);
    v = fmaf(c[2 * j], w[3], v);
    v = fmaf(c[2 * j + 1], w[4], v);
    o[j] = fmaf(c[2 * j + 2], w[5], v);
  }
}
__device__ __forceinline__ void dw_row_s2_d(const float (&d)[16], const float (&w)[9], const float (&o)[7], uint32_t ob) {
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    float v = fmaf(d[2 * j], w[6], o[j]);
    v = fmaf(d[2 * j + 1], w[7], v);
    v = fma_sat(d[2 * j + 2], w[8], v);
    st_half((ob ^ static_cast<uint32_t>((j & 7) << 4)) + j * 128, v);
  }
}

template <int S, int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
mbconv_rows_kernel(const __grid_constant__ MrTensorMaps maps, const __grid_constant__ MrParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* s_x = smem;
  uint8_t* s_w1 = smem + p.off_w1;
  uint8_t* s_out = smem + p.off_out;
  MrCtrl* ctrl = reinterpret_cast<MrCtrl*>(smem + p.off_ctrl);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_dw_warps = 4 * p.lay.WQ;
  const int nch = p.lay.nchunks;
  const int NSc = kSlots / nch;                           // E ring slots per chunk (6, 3 or 2)
  const int RPI = 64 / p.RP;                              // image rows per slot: 1, 2 or 4
  const int g_log2 = p.G == 2 ? 1 : (p.G == 4 ? 2 : 3);

  // contiguous range of frame segments of this CTA
  const int units = p.N * p.segs;
  const int per = units / static_cast<int>(gridDim.x), rem = units % static_cast<int>(gridDim.x);
  const int bx = static_cast<int>(blockIdx.x);
  const int u0 = bx * per + (bx < rem ? bx : rem);
  const int n_units = per + (bx < rem ? 1 : 0);
  const int flush = (S == 1 && n_units > 0) ? 1 : 0;      // stride 1: one more step closes the last frame's last row
  const int T = n_units * p.SPF + flush;                  // steps of this CTA
  const int spi_log2 = p.SPI - 1;                         // SPI is 1 or 2
  const int TS = (T + p.SPI - 1) >> spi_log2;             // project items (accumulator tiles)

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) {
      mbar_init(&ctrl->x_full[i], 1);
      mbar_init(&ctrl->x_empty[i], nch > 1 ? 2 : 1);      // one commit per expand issuer warp
    }
    mbar_init(&ctrl->w_full, 1);
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&ctrl->e_full[i], 1);
      mbar_init(&ctrl->e_free[i], static_cast<uint32_t>(p.lay.warps[i % nch]));
    }
    for (int i = 0; i < kMrMaxBufs; ++i) {
      mbar_init(&ctrl->a2_full[i], static_cast<uint32_t>(i < p.NB ? p.a2_cnt[i] : 1));
      mbar_init(&ctrl->a2_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctrl->d2_full[i], 1);
      mbar_init(&ctrl->d2_free[i], 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.w1);
    tma_prefetch_desc(&maps.w2);
    tma_prefetch_desc(&maps.out);
    if (S == 1) {
      tma_prefetch_desc(&maps.out_rest);
      tma_prefetch_desc(&maps.out_one);
    }
  }
  if (warp == 1) {
    tmem_alloc(&ctrl->tmem_base, 512);
    tmem_relinquish();
  }
  if (warp >= kFirstDwWarp) {
    // the A2 operand buffers start as zeros: K columns no lane writes multiply zero weights and must hold finite values
    const int ct = threadIdx.x - kFirstDwWarp * 32;
    const int a2_total = p.off_out - p.off_a2;
    for (int i = ct; i < a2_total / 16; i += n_dw_warps * 32)
      reinterpret_cast<uint4*>(smem + p.off_a2)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  pdl_launch_dependents();

  if (warp == 0) {
    // ============================ TMA producer: weights once, then G input rows per step ============================
    uint32_t w_bytes = static_cast<uint32_t>(nch) * 16384u;
    for (int c = 0; c < nch; ++c) w_bytes += static_cast<uint32_t>(p.lay.a2_bytes[c] >> 14) * static_cast<uint32_t>(p.cout_pad) * 128u;
    mbar_arrive_expect_tx_elect(&ctrl->w_full, w_bytes);
    for (int c = 0; c < nch; ++c) tma_load_2d_elect(s_w1 + c * 16384, &maps.w1, &ctrl->w_full, 0, c * 128);
    for (int c = 0; c < nch; ++c)
      for (int sub = 0; sub < (p.lay.a2_bytes[c] >> 14); ++sub)
        tma_load_2d_elect(smem + p.w2_off[c] + sub * p.cout_pad * 128, &maps.w2, &ctrl->w_full, c * 128 + sub * 64, 0);
    pdl_wait_prior_grid();
    int stage = 0, k = 0, unit = u0;
    uint32_t ph = 0;
    for (int t = 0; t < T; ++t) {
      while (!mbar_try_wait(&ctrl->x_empty[stage], ph ^ 1u)) __nanosleep(64);
      const bool is_flush = flush && t == T - 1;
      const int n = is_flush ? p.N : unit >> (p.segs - 1);   // the closing step reads past the batch: all zeros
      const int seg = is_flush ? 0 : unit & (p.segs - 1);    // segs is 1 or 2
      mbar_arrive_expect_tx_elect(&ctrl->x_full[stage], static_cast<uint32_t>(p.x_stage));
      tma_load_4d_elect(s_x + stage * p.x_stage, &maps.x, &ctrl->x_full[stage], 0, seg * p.OWseg * S - 1, k * p.G, n);
      if (++stage == p.XS) {
        stage = 0;
        ph ^= 1u;
      }
      if (++k == p.SPF) {
        k = 0;
        ++unit;
      }
    }
  } else if (warp == 1 || (warp == 3 && nch > 1)) {
    // ============================ expand MMA issuers: one GEMM per (64 pixels = RPI image rows, chunk) ============================
    // E[slot] (128 lanes x 64 columns) = W1[chunk] * X[rows]^T; chunk c owns slots c, c + nch, ..: NSc slots of every
    // chunk are in flight, so a GEMM is issued while the depthwise warps still work on the slots before it.
    // Warp 1 issues chunks 0 and 2, warp 3 chunk 1: a single issuing warp spends ~500 cycles of dependent instruction
    // latency per GEMM next to the busy depthwise warps of its scheduler, which bounded the first version.
    const int c_first = warp == 1 ? 0 : 1, c_step = nch > 1 ? 2 : 1;
    const uint32_t idesc1 = make_idesc_f16_f32(128, 64);
    const int IPS = p.G / RPI, XS = p.XS;                                    // items per step
    const uint32_t k1 = static_cast<uint32_t>(p.k1steps);
    const uint32_t item_lo = static_cast<uint32_t>(64 * 128) >> 4;           // descriptor step between items of a stage
    const uint32_t x_lo0 = smem_desc_lo(smem_u32(s_x));
    const uint32_t w_lo0 = smem_desc_lo(smem_u32(s_w1));
    const uint32_t full0 = smem_u32(&ctrl->e_full[0]), free0 = smem_u32(&ctrl->e_free[0]);
    const uint32_t ring_slots = static_cast<uint32_t>(NSc * nch);
    wait_sleep(&ctrl->w_full, 0);
    tc_fence_after();
    int stage = 0;
    uint32_t row_slot = 0;          // first slot (chunk 0) of the current ring position
    uint32_t ph = 0, fph = 0;       // fph: parity that says "the previous use of this slot has been released"
    bool first_lap = true;
    // the loop body per K-step count: the issuing warp's instruction stream is the critical resource
    auto run = [&](auto kc) {
      constexpr int K1 = decltype(kc)::value;
      for (int t = 0; t < T; ++t) {
        wait_sleep(&ctrl->x_full[stage], ph);
        uint32_t lb = x_lo0 + static_cast<uint32_t>(stage) * (static_cast<uint32_t>(p.x_stage) >> 4);
        for (int g = 0; g < IPS; ++g, lb += item_lo) {
          for (int c = c_first; c < nch; c += c_step) {
            const uint32_t slot = row_slot + static_cast<uint32_t>(c);
            if (!first_lap) {
              const uint32_t fb = free0 + slot * 8u;
              uint32_t done;
              do {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                    : "=r"(done)
                    : "r"(fb), "r"(fph)
                    : "memory");
              } while (!done);
            }
            tc_fence_after();
            umma_group_commit_elect<K1>(tmem_base + slot * 64u, w_lo0 + static_cast<uint32_t>(c) * (16384u >> 4), lb, idesc1,
                                        full0 + slot * 8u);
          }
          row_slot += static_cast<uint32_t>(nch);
          if (row_slot == ring_slots) {
            row_slot = 0;
            if (first_lap) first_lap = false;
            else fph ^= 1u;
          }
        }
        umma_commit_elect(&ctrl->x_empty[stage]);
        if (++stage == XS) {
          stage = 0;
          ph ^= 1u;
        }
      }
    };
    if (k1 == 1) run(std::integral_constant<int, 1>{});
    else if (k1 == 2) run(std::integral_constant<int, 2>{});
    else if (k1 == 3) run(std::integral_constant<int, 3>{});
    else run(std::integral_constant<int, 4>{});
  } else if (warp == 2) {
    // ============================ project MMA issuer: D2[item] = sum over chunks A2[chunk] * W2[chunk]^T ============================
    const uint32_t idesc2 = make_idesc_f16_f32(128, static_cast<uint32_t>(p.cout_pad));
    wait_sleep(&ctrl->w_full, 0);
    tc_fence_after();
    for (int ts = 0; ts < TS; ++ts) {
      const int dbuf = ts & 1;
      const uint32_t d2 = tmem_base + static_cast<uint32_t>(kD2Col + dbuf * 64);
      for (int c = 0; c < nch; ++c) {
        const int buf = c * 2 + (ts & 1), use = ts >> 1;
        if (c == 0) wait_sleep(&ctrl->d2_free[dbuf], ((static_cast<uint32_t>(ts) >> 1) & 1u) ^ 1u);
        wait_sleep(&ctrl->a2_full[buf], static_cast<uint32_t>(use) & 1u);
        tc_fence_after();
        const uint32_t la = smem_desc_lo(smem_u32(smem + p.a2_off[buf]));
        const uint32_t lb = smem_desc_lo(smem_u32(smem + p.w2_off[c]));
        const uint32_t sub_b = static_cast<uint32_t>(p.cout_pad) * 128u;
        for (int ks = 0; ks < p.lay.ksteps[c]; ++ks) {
          const uint32_t sub = static_cast<uint32_t>(ks >> 2), kk = static_cast<uint32_t>(ks & 3);
          umma_f16_ss_lo_elect(d2, la + ((sub * static_cast<uint32_t>(p.a2_sub)) >> 4) + kk * 2u, lb + ((sub * sub_b) >> 4) + kk * 2u, idesc2,
                               (c | ks) != 0 ? 1u : 0u);
        }
        umma_commit_elect(&ctrl->a2_free[buf]);
        if (c == nch - 1) umma_commit_elect(&ctrl->d2_full[dbuf]);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ============================ epilogue: D2 -> +bias (+ residual) -> fp16 -> staging -> TMA store ============================
    const int quarter = warp & 3;
    const int et = threadIdx.x - 128;
    const int prow = quarter * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    // decode the accumulator row: (step half, output row inside the step, strip, column inside the strip)
    const int rows_per_step = 128 >> spi_log2;
    const int half = prow / rows_per_step, pp = prow - half * rows_per_step;
    const int oi = pp / p.RP, os = (pp % p.RP) >> 4, oj = pp & 15;
    const bool lane_valid = oj < p.OW && os < p.SPR && oi < p.OR;
    const int srow = (half * p.OR + oi) * p.OWseg + os * p.OW + oj;     // row of the dense staging tile
    // stride 1, first step of a frame: rows 1..G-1 move up by one row, row 0 goes behind them at a 1024-B boundary
    const int first_base = ((p.G - 1) * p.OWseg + 7) & ~7;
    const int srow_first = oi == 0 ? first_base + os * p.OW + oj : srow - p.OWseg;
    const int ox_seg = os * p.OW + oj;
    const int ncg = p.cout_pad >> 4;
    pdl_wait_prior_grid();
    for (int ts = 0; ts < TS; ++ts) {
      const int t0 = ts * p.SPI;
      const int unit = u0 + t0 / p.SPF, k0 = t0 % p.SPF;
      const int n = unit >> (p.segs - 1), seg = unit & (p.segs - 1);
      const bool is_flush = flush && t0 == T - 1;
      const int dbuf = ts & 1;
      wait_spin(&ctrl->d2_full[dbuf], (static_cast<uint32_t>(ts) >> 1) & 1u);
      tc_fence_after();
      if (et == 0) tma_store_wait_read0();     // the previous item's stores have released the staging tile
      named_barrier_sync(kEpiBarrier, 128);
      // output pixel of this accumulator row
      const int pn = (unit - 1) >> (p.segs - 1), pseg = (unit - 1) & (p.segs - 1);   // the frame segment before this one
      int on = n, oseg = seg, oy;
      if (S == 1) {
        oy = k0 * p.G - 1 + oi;
        if (oy < 0) {              // first row of a segment's first step: the last row of the previous segment
          on = unit > 0 ? pn : -1;
          oseg = pseg;
          oy = p.SPF * p.G - 1;
        }
      } else {
        oy = (k0 + half) * p.OR + oi;
      }
      const __half* rp = nullptr;
      if (S == 1 && lane_valid && p.residual != nullptr && on >= 0 && on < p.N && oy < p.H)
        rp = p.residual + ((static_cast<long long>(on) * p.H + oy) * p.W + oseg * p.OWseg + ox_seg) * p.res_stride;
      const int sr = (S == 1 && k0 == 0) ? srow_first : srow;
      uint8_t* srow_p = s_out + sr * 128;
      for (int cg = 0; cg < ncg; ++cg) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + lane_sel + static_cast<uint32_t>(kD2Col + dbuf * 64 + cg * 16), v);
        tmem_ld_wait();
        if (cg == ncg - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctrl->d2_free[dbuf]);
        }
        if (lane_valid) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int co = cg * 16 + h * 8;
            const float4 ba = __ldg(reinterpret_cast<const float4*>(p.bias3 + co));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias3 + co + 4));
            float x[8] = {__uint_as_float(v[h * 8 + 0]) + ba.x, __uint_as_float(v[h * 8 + 1]) + ba.y,
                          __uint_as_float(v[h * 8 + 2]) + ba.z, __uint_as_float(v[h * 8 + 3]) + ba.w,
                          __uint_as_float(v[h * 8 + 4]) + bb.x, __uint_as_float(v[h * 8 + 5]) + bb.y,
                          __uint_as_float(v[h * 8 + 6]) + bb.z, __uint_as_float(v[h * 8 + 7]) + bb.w};
            if (rp != nullptr && co + 8 <= p.Cout) {
              const uint4 rv = __ldg(reinterpret_cast<const uint4*>(rp + co));
              const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(rh[j]);
                x[2 * j] += f.x;
                x[2 * j + 1] += f.y;
              }
            }
            uint4 ov;
            __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
            for (int j = 0; j < 4; ++j) oh2[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
            *reinterpret_cast<uint4*>(srow_p + (((cg * 2 + h) ^ (sr & 7)) << 4)) = ov;
          }
        }
      }
      fence_proxy_async();
      named_barrier_sync(kEpiBarrier, 128);
      if (et == 0) {
        if (S == 1) {
          if (k0 == 0) {
            if (ts > 0 && p.SPF * p.G - 1 < p.H)
              tma_store_4d(&maps.out_one, s_out + first_base * 128, 0, pseg * p.OWseg, p.SPF * p.G - 1, pn);
            if (!is_flush) tma_store_4d(&maps.out_rest, s_out, 0, seg * p.OWseg, 0, n);
          } else {
            tma_store_4d(&maps.out, s_out, 0, seg * p.OWseg, k0 * p.G - 1, n);
          }
        } else {
          tma_store_4d(&maps.out, s_out, 0, seg * p.OWseg, k0 * p.OR, n);
        }
        tma_store_commit();
      }
    }
    if (et == 0) tma_store_wait_all();
  } else if (warp >= kFirstDwWarp && warp < kFirstDwWarp + n_dw_warps) {
    // ============================ depthwise warps: E (TMEM) -> 3x3 -> A2 ============================
    const int quarter = warp & 3, task = (warp - kFirstDwWarp) >> 2;
    const int chunk = p.lay.task_chunk[quarter][task], strip = p.lay.task_strip[quarter][task];
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    const int l128 = quarter * 32 + lane;
    float w[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) w[i] = __ldg(p.dwp + (chunk * 11 + i) * 128 + l128);
    const float b1 = __ldg(p.dwp + (chunk * 11 + 9) * 128 + l128);
    const float b2 = __ldg(p.dwp + (chunk * 11 + 10) * 128 + l128);
    const int kpos = p.lay.lane_kpos[chunk][l128];
    // byte offset of this lane's K column inside an A2 buffer (bits 4-6 = 16-byte chunk, XORed with pixel & 7 per store)
    const uint32_t kcol = static_cast<uint32_t>((kpos >> 6) * p.a2_sub + (((kpos & 63) >> 3) << 4) + (kpos & 7) * 2);
    const uint32_t e_col0 = tmem_base + lane_sel + static_cast<uint32_t>(strip * 14);
    const uint32_t smem_base = smem_u32(smem);
    const bool last_strip = strip == p.SPR - 1;

    // per-step state (set by the first row of a step)
    int k = 0, unit = u0;
    uint32_t a2_base = 0;
    int a2_buf = 0;
    bool step_flush = false, zero_first = false, zero_last_col = false;

    // E ring position of this warp's chunk: slot = chunk + nch * e_j, parity e_ph
    int e_j = 0;
    uint32_t e_ph = 0;
    auto step_begin = [&](int t) {
      step_flush = flush && t == T - 1;
      zero_first = strip == 0 && (unit & (p.segs - 1)) == 0;
      zero_last_col = last_strip && (unit & (p.segs - 1)) == p.segs - 1;
      const int ts = t >> spi_log2, sub = t & (p.SPI - 1);
      a2_buf = chunk * 2 + (ts & 1);
      a2_base = smem_base + static_cast<uint32_t>(p.a2_off[a2_buf]) +
                static_cast<uint32_t>((sub * 64 + strip * 16) * 128) + kcol;
      if (sub == 0) wait_spin(&ctrl->a2_free[a2_buf], ((static_cast<uint32_t>(ts) >> 1) & 1u) ^ 1u);
    };
    // next input row of this warp's chunk: wait for its GEMM (first row of a slot), read the strip, hand the slot back
    // after its last row
    int e_ri = 0;
    auto next_row = [&](bool zero_last, bool zero_row, float (&r)[16]) {
      const int slot = chunk + nch * e_j;
      if (e_ri == 0) {
        wait_spin(&ctrl->e_full[slot], e_ph);
        tc_fence_after();
      }
      load_row(e_col0 + static_cast<uint32_t>(slot * 64 + e_ri * p.RP), b1, zero_first, zero_last, zero_row, r);
      if (++e_ri == RPI) {
        e_ri = 0;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->e_free[slot]);
        if (++e_j == NSc) {
          e_j = 0;
          e_ph ^= 1u;
        }
      }
    };
    auto step_end = [&]() {
      fence_proxy_async();                  // A2 writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->a2_full[a2_buf]);
      if (++k == p.SPF) {
        k = 0;
        ++unit;
      }
    };

    if (S == 1) {
      // three input rows roll through ra / rb / rc; every new row completes one output row (the one above it)
      float ra[16], rb[16], rc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) ra[i] = rb[i] = 0.f;
      const int R = T << g_log2;
      auto do_row = [&](int rr, float (&r0)[16], float (&r1)[16], float (&r2)[16]) {
        const int pi = rr & (p.G - 1);
        if (pi == 0) step_begin(rr >> g_log2);
        const int y = step_flush ? p.H : k * p.G + pi;
        next_row(zero_last_col, y >= p.H, r2);
        const uint32_t ob = a2_base + static_cast<uint32_t>(pi * p.RP * 128);
        if (k == 0 && pi < 2) {
          if (pi == 0) dw_row_s1<1>(r0, r1, r2, w, b2, ob);
          else dw_row_s1<2>(r0, r1, r2, w, b2, ob);
        } else {
          dw_row_s1<0>(r0, r1, r2, w, b2, ob);
        }
        if (pi == p.G - 1) step_end();
      };
      for (int rr = 0; rr < R; rr += 3) {
        do_row(rr, ra, rb, rc);
        if (rr + 1 < R) do_row(rr + 1, rb, rc, ra);
        if (rr + 2 < R) do_row(rr + 2, rc, ra, rb);
      }
    } else {
      float ra[16], rc[16], rd[16];
      const int pps_log2 = g_log2 - 1, PPS = 1 << pps_log2;     // row pairs per step
      const int P = T << pps_log2;
      auto do_pair = [&](int pp, float (&a)[16], float (&c)[16], float (&d)[16]) {
        const int pi = pp & (PPS - 1);
        if (pi == 0) step_begin(pp >> pps_log2);
        if (k == 0 && pi == 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) a[i] = 0.f;        // input row -1 of a new frame
        }
        float o[7];
        next_row(false, false, c);
        dw_row_s2_ac(a, c, w, b2, o);
        next_row(false, false, d);
        dw_row_s2_d(d, w, o, a2_base + static_cast<uint32_t>(pi * p.RP * 128));
        if (pi == PPS - 1) step_end();
      };
      for (int pp = 0; pp < P; pp += 2) {
        do_pair(pp, ra, rc, rd);
        if (pp + 1 < P) do_pair(pp + 1, rd, rc, ra);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool mbrows_layout(int cexp, int spr, MrLayout* L) {
  memset(L, 0, sizeof(*L));
  if (cexp < 16 || cexp % 16 != 0 || (spr != 1 && spr != 2 && spr != 4)) return false;
  for (int c = 0; c < kMrMaxChunks; ++c)
    for (int l = 0; l < 128; ++l) {
      L->lane_ch[c][l] = -1;
      L->lane_kpos[c][l] = static_cast<int16_t>(l);
    }
  int tq[4] = {0, 0, 0, 0};
  bool ok = true;
  auto add_task = [&](int q, int chunk, int strip) {
    if (tq[q] >= kMrMaxWQ) {
      ok = false;
      return;
    }
    L->task_chunk[q][tq[q]] = static_cast<int8_t>(chunk);
    L->task_strip[q][tq[q]] = static_cast<int8_t>(strip);
    ++tq[q];
    ++L->warps[chunk];
  };
  int groups = cexp / 32, ch = 0, nc = 0;
  const bool half = (cexp % 32) == 16;
  // full chunks: one 32-channel group per lane quarter, every quarter walks every strip
  while (groups >= 4) {
    if (nc >= kMrMaxChunks) return false;
    for (int l = 0; l < 128; ++l) {
      L->lane_ch[nc][l] = static_cast<int16_t>(ch + l);
      L->lane_kpos[nc][l] = static_cast<int16_t>(l);
    }
    L->ksteps[nc] = 8;
    L->a2_bytes[nc] = 32768;
    for (int q = 0; q < 4; ++q)
      for (int s = 0; s < spr; ++s) add_task(q, nc, s);
    groups -= 4;
    ch += 128;
    ++nc;
  }
  // two groups left: replicated as [A B A B]; quarters 0-1 take the first half of the strips, 2-3 the second half
  if (groups >= 2) {
    if (spr < 2 || nc >= kMrMaxChunks) return false;
    for (int q = 0; q < 4; ++q)
      for (int l = 0; l < 32; ++l) {
        L->lane_ch[nc][q * 32 + l] = static_cast<int16_t>(ch + (q & 1) * 32 + l);
        L->lane_kpos[nc][q * 32 + l] = static_cast<int16_t>((q & 1) * 32 + l);
      }
    L->ksteps[nc] = 4;
    L->a2_bytes[nc] = 16384;
    for (int q = 0; q < 4; ++q)
      for (int s = 0; s < spr / 2; ++s) add_task(q, nc, (q >> 1) * (spr / 2) + s);
    groups -= 2;
    ch += 64;
    ++nc;
  }
  // one group (32 channels) or half a group (16) left: replicated into every quarter, quarter q takes strip q
  // (four strips), or split over quarter pairs (two strips)
  for (int pass = 0; pass < 2; ++pass) {
    const int cnt = pass == 0 ? (groups >= 1 ? 32 : 0) : (half ? 16 : 0);
    if (cnt == 0) continue;
    if (nc >= kMrMaxChunks || spr == 1) return false;
    const int split = spr == 4 ? 1 : 2;            // channel sub-groups side by side in different quarters
    const int per = cnt / split;
    for (int q = 0; q < 4; ++q)
      for (int l = 0; l < 32; ++l) {
        const int sub = split == 1 ? 0 : (q & 1);
        if (l < per) {
          L->lane_ch[nc][q * 32 + l] = static_cast<int16_t>(ch + sub * per + l);
          L->lane_kpos[nc][q * 32 + l] = static_cast<int16_t>(sub * per + l);
        } else {
          L->lane_kpos[nc][q * 32 + l] = static_cast<int16_t>(cnt + sub * (32 - per) + (l - per));
        }
      }
    L->ksteps[nc] = (cnt + 15) / 16;
    L->a2_bytes[nc] = 16384;
    for (int q = 0; q < 4; ++q) add_task(q, nc, spr == 4 ? q : (q >> 1));
    if (pass == 0) groups -= 1;
    ch += cnt;
    ++nc;
  }
  if (!ok || groups != 0 || ch != cexp || nc == 0) return false;
  if (tq[0] != tq[1] || tq[0] != tq[2] || tq[0] != tq[3] || tq[0] < 1) return false;
  L->nchunks = nc;
  L->WQ = tq[0];
  return true;
}

bool mbrows_plan(MrParams* p) {
  if (p->S != 1 && p->S != 2) return false;
  if (p->Cin < 8 || p->Cin > 64 || p->Cin % 8 != 0) return false;
  if (p->Cout < 8 || p->Cout > 64 || p->Cout % 8 != 0) return false;
  if (p->N < 1 || p->H < 2 || p->W < 2) return false;
  p->Ho = (p->H - 1) / p->S + 1;
  p->Wo = (p->W - 1) / p->S + 1;
  p->OW = p->S == 1 ? 14 : 7;
  p->segs = 1;
  int wo_seg = p->Wo;
  if (p->S == 2) {
    if (p->W % 2 != 0) return false;
    if (p->Wo == 56) {
      p->segs = 2;
      wo_seg = 28;
    }
  } else if (p->Wo == 112) {
    p->segs = 2;                 // two 56-pixel segments per row, each with its own halo columns
    wo_seg = 56;
  }
  if (wo_seg % p->OW != 0) return false;
  p->SPR = wo_seg / p->OW;
  if (p->SPR != 1 && p->SPR != 2 && p->SPR != 4) return false;
  p->RP = p->SPR * 16;
  p->OWseg = wo_seg;
  p->k1steps = (p->Cin + 15) / 16;
  p->cout_pad = (p->Cout + 15) / 16 * 16;
  if (!mbrows_layout(p->Cexp, p->SPR, &p->lay)) return false;
  const int nc = p->lay.nchunks;
  // Steps of 128 pixels (G = 128 / RP rows); when the two A2 buffers per chunk do not fit next to the weights (three
  // chunks at 14 x 14), steps of 64 pixels: the project GEMM then reads 128-row operand tiles whose upper half is
  // whatever follows in shared memory -- rows of M do not mix, those accumulator rows are never stored.
  // (One ring of buffers shared by all chunks would serialise the chunks inside a step: the warps of the last chunk
  // could only start once the project GEMM of the first one has drained its buffer.)
  for (int attempt = 0; attempt < 4; ++attempt) {
    const int xs = (attempt & 1) == 0 ? 3 : 2;
    p->G = (attempt < 2 ? 128 : 64) / p->RP;
    if (p->G < 64 / p->RP || p->G < 2) continue;
    if (attempt >= 2 && p->S == 2) continue;
    p->OR = p->S == 1 ? p->G : p->G / 2;
    p->SPF = (p->H + p->G - 1) / p->G;
    p->SPI = (p->S == 2 && p->OR * p->RP == 64 && p->SPF % 2 == 0) ? 2 : 1;
    p->x_stage = p->G * p->RP * 128;
    p->a2_sub = (p->S == 1 ? p->G * p->RP : 128) * 128;
    int off = xs * p->x_stage;
    off = (off + 1023) & ~1023;
    p->off_w1 = off;
    off += nc * 16384;
    p->off_w2 = off;
    for (int c = 0; c < nc; ++c) {
      p->w2_off[c] = off;
      off += (p->lay.a2_bytes[c] >> 14) * p->cout_pad * 128;
    }
    off = (off + 1023) & ~1023;
    p->off_a2 = off;
    p->NB = 2 * nc;
    for (int c = 0; c < nc; ++c)
      for (int par = 0; par < 2; ++par) {
        p->a2_off[c * 2 + par] = off;
        p->a2_cnt[c * 2 + par] = p->lay.warps[c] * p->SPI;
        off += (p->lay.a2_bytes[c] >> 14) * p->a2_sub;
      }
    p->off_out = off;
    off += 16384;
    p->off_ctrl = off;
    off += 1024;
    p->smem = off;
    p->XS = xs;
    if (off <= 227 * 1024) return true;
  }
  return false;
}

cudaError_t launch_mbconv_rows(const MrTensorMaps& maps, const MrParams& p, int sm_count, cudaStream_t stream) {
  static_assert(sizeof(MrCtrl) <= 1024, "ctrl block too large");
  using Kern = void (*)(const MrTensorMaps, const MrParams);
  const int wq = p.lay.WQ;
  Kern kern = nullptr;
  int threads = 0;
  if (wq <= 3) {
    kern = p.S == 1 ? mbconv_rows_kernel<1, 640> : mbconv_rows_kernel<2, 640>;
    threads = 640;
  } else if (wq == 4) {
    kern = p.S == 1 ? mbconv_rows_kernel<1, 768> : mbconv_rows_kernel<2, 768>;
    threads = 768;
  } else {
    kern = p.S == 1 ? mbconv_rows_kernel<1, 896> : mbconv_rows_kernel<2, 896>;
    threads = 896;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  const int units = p.N * p.segs;
  int grid = units < sm_count ? units : sm_count;
  if (grid < 1) grid = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = getenv("AF_NO_PDL") == nullptr;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, maps, p);
}

}  // namespace af
