// Fused inverted-residual block with the depthwise 3x3 on the tensor core (see mbconv_tc.cuh).
#include "mbconv_tc.cuh"

#include <cstdio>
#include <cstdlib>

#include "ptx.cuh"

namespace af {

using namespace ptx;

namespace {

constexpr int kFirstEpiWarp = 3;      // warps 0-2: TMA producer, expand MMA issuer, depthwise + project MMA issuer
constexpr int kWbTileBytes = 2048;    // one tap's block-diagonal B operand: 16 rows x 128 B (K = 64), 128-B swizzled
constexpr int kWbBufBytes = 9 * kWbTileBytes;
constexpr int kProjBarrier = 1;       // named barrier of the depthwise-side warps (output staging hand-over)

struct __align__(8) Mb3Ctrl {
  uint64_t x_full[2], x_empty[2];
  uint64_t w_full;
  uint64_t d1_full[2], d1_empty[2];
  uint64_t e_full[2], e_empty[2];
  uint64_t wb_empty[2];
  uint64_t d3_full[2], d3_empty[2];
  uint64_t a2_full[2], a2_empty[2];
  uint64_t d2_full[2], d2_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void wait_backoff(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(32);
}

// 32 accumulator columns of one row -> +bias, ReLU6 -> fp16 -> four 16-byte chunks of a 128-byte swizzled row
__device__ __forceinline__ void bias_relu6_store(const uint32_t (&v)[32], const float* __restrict__ b, uint8_t* rowp,
                                                 int row, int chunk0, bool zero) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 ba = *reinterpret_cast<const float4*>(b + i * 8);
    const float4 bb = *reinterpret_cast<const float4*>(b + i * 8 + 4);
    const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
    uint4 ov;
    __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __half2 h = __floats2half2_rn(__uint_as_float(v[i * 8 + 2 * j]) + bv[2 * j],
                                    __uint_as_float(v[i * 8 + 2 * j + 1]) + bv[2 * j + 1]);
      oh2[j] = __hmin2(__hmax2(h, __float2half2_rn(0.f)), __float2half2_rn(6.f));
    }
    if (zero) ov = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(rowp + (((chunk0 + i) ^ (row & 7)) << 4)) = ov;
  }
}

template <int S>
__global__ void __launch_bounds__(kMb3Threads, 1)
mbconv3_kernel(const __grid_constant__ MbTensorMaps maps, const Mb3Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* s_x = smem;
  uint8_t* s_w1 = smem + p.off_w1;
  uint8_t* s_w2 = smem + p.off_w2;
  uint8_t* s_a2 = smem + p.off_a2;
  uint8_t* s_out = smem + p.off_out;
  uint8_t* s_e = smem + p.off_e;
  uint8_t* s_wb = smem + p.off_wb;
  const int CE = p.nc * 64;
  __half* s_dwh = reinterpret_cast<__half*>(smem + p.off_c);            // [9][CE] fp16 depthwise weights
  float* s_b1 = reinterpret_cast<float*>(smem + p.off_c + 18 * CE);      // [CE]
  float* s_b2 = s_b1 + CE;                                               // [CE]
  float* s_b3 = s_b2 + CE;                                               // [64]
  Mb3Ctrl* ctrl = reinterpret_cast<Mb3Ctrl*>(smem + p.off_ctrl);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x_buf_bytes = p.Mtiles * 16384;
  const int e_buf_bytes = p.e_rows * 128;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int n_tiles = tiles_per_img * p.N;
  const uint32_t w2_chunk = static_cast<uint32_t>(p.cout_pad) * 128u;
  const int d1_cols = p.Mtiles * 64;
  const int d3_col = p.Mtiles <= 2 ? 256 : 384;   // TMEM: D1 2 x Mtiles x 64 | D3 D3B x 64 | D2 D2B x 64  (<= 512 columns)
  const int d2_col = p.Mtiles <= 2 ? 384 : 448;
  const int nA = p.nA, nB = 4 - p.nA;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctrl->x_full[i], 1);
      mbar_init(&ctrl->x_empty[i], 1);
      mbar_init(&ctrl->d1_full[i], 1);
      mbar_init(&ctrl->d1_empty[i], 4 * nA);
      mbar_init(&ctrl->e_full[i], 4 * nA);
      mbar_init(&ctrl->e_empty[i], 1);
      mbar_init(&ctrl->wb_empty[i], 1);
      mbar_init(&ctrl->d3_full[i], 1);
      mbar_init(&ctrl->d3_empty[i], 4 * nB);
      mbar_init(&ctrl->a2_full[i], 4 * nB);
      mbar_init(&ctrl->a2_empty[i], 1);
      mbar_init(&ctrl->d2_full[i], 1);
      mbar_init(&ctrl->d2_empty[i], 4 * nB);
    }
    mbar_init(&ctrl->w_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&maps.x);
    tma_prefetch_desc(&maps.w1);
    tma_prefetch_desc(&maps.w2);
    tma_prefetch_desc(&maps.out);
  }
  if (warp == 1) {
    tmem_alloc(&ctrl->tmem_base, 512);
    tmem_relinquish();
  }
  if (warp >= kFirstEpiWarp) {
    // launch constants -> shared memory; the block-diagonal weight tiles start as zeros (only their diagonals are
    // rewritten per chunk)
    const int ct = threadIdx.x - kFirstEpiWarp * 32;
    constexpr int nt = kMb3Threads - kFirstEpiWarp * 32;
    for (int i = ct; i < p.WB * kWbBufBytes / 16; i += nt) reinterpret_cast<uint4*>(s_wb)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = ct; i < 9 * CE; i += nt) s_dwh[i] = __float2half_rn(p.dw_w[i]);
    for (int i = ct; i < CE; i += nt) {
      s_b1[i] = p.bias1[i];
      s_b2[i] = p.bias2[i];
    }
    if (ct < 64) s_b3[ct] = ct < p.cout_pad ? p.bias3[ct] : 0.f;
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_base;
  pdl_launch_dependents();

  const int my_tiles = static_cast<int>(blockIdx.x) < n_tiles
                           ? (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                 static_cast<int>(gridDim.x)
                           : 0;
  const int G = my_tiles * p.nc;   // channel chunks this CTA goes through, across all its tiles

  if (warp == 0) {
    // ============================ TMA producer: weights once, then one input window per tile ============================
    if (lane == 0) {
      mbar_arrive_expect_tx(&ctrl->w_full, static_cast<uint32_t>(p.nc) * (8192u + w2_chunk));
      for (int c = 0; c < p.nc; ++c) tma_load_2d(s_w1 + c * 8192, &maps.w1, &ctrl->w_full, 0, c * 64);
      for (int c = 0; c < p.nc; ++c) tma_load_2d(s_w2 + c * w2_chunk, &maps.w2, &ctrl->w_full, c * 64, 0);
      pdl_wait_prior_grid();
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int xb = it % p.XB;
        const uint32_t ph = static_cast<uint32_t>(it / p.XB) & 1u;
        while (!mbar_try_wait(&ctrl->x_empty[xb], ph ^ 1u)) __nanosleep(128);
        const int n = tile / tiles_per_img;
        const int r = tile - n * tiles_per_img;
        const int th_i = r / p.tiles_w, tw_i = r - th_i * p.tiles_w;
        mbar_arrive_expect_tx(&ctrl->x_full[xb], static_cast<uint32_t>(p.n_rows) * 128u);
        tma_load_4d(s_x + xb * x_buf_bytes, &maps.x, &ctrl->x_full[xb], 0, tw_i * p.TW * S - 1, th_i * p.TH * S - 1, n);
      }
    }
  } else if (warp == 1) {
    // ============================ expand MMA issuer: D1[g & 1] = X * W1[chunk]^T ============================
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc_f16_f32(128, 64);
      wait_backoff(&ctrl->w_full, 0);
      tc_fence_after();
      int it = 0, c = 0;
      for (int g = 0; g < G; ++g) {
        const int xb = it % p.XB;
        if (c == 0) wait_backoff(&ctrl->x_full[xb], static_cast<uint32_t>(it / p.XB) & 1u);
        wait_backoff(&ctrl->d1_empty[g & 1], ((static_cast<uint32_t>(g) >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t xa = smem_u32(s_x + xb * x_buf_bytes);
        const uint64_t db = make_smem_desc_sw128(smem_u32(s_w1 + c * 8192));
        for (int m = 0; m < p.Mtiles; ++m) {
          const uint64_t da = make_smem_desc_sw128(xa + static_cast<uint32_t>(m) * 16384u);
          const uint32_t d = tmem_base + static_cast<uint32_t>((g & 1) * d1_cols + m * 64);
          for (int k = 0; k < p.k1steps; ++k)
            umma_f16_ss(d, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc1, k != 0 ? 1u : 0u);
        }
        umma_commit(&ctrl->d1_full[g & 1]);
        if (c == p.nc - 1) umma_commit(&ctrl->x_empty[xb]);   // the tile's input window has been consumed
        if (++c == p.nc) {
          c = 0;
          ++it;
        }
      }
    }
  } else if (warp == 2) {
    // ============================ depthwise + project MMA issuer ============================
    // Iteration g: rebuild the diagonal weight tiles of chunk g (all lanes), issue the depthwise MMAs of chunk g, then
    // the project MMAs of chunk g-1 (whose A2 tile the depthwise-epilogue warps finish meanwhile).
    const uint32_t idesc_dw = make_idesc_f16_f32(128, 16);
    const uint32_t idesc2 = make_idesc_f16_f32(128, static_cast<uint32_t>(p.cout_pad));
    int tap_off[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int kh = t / 3, kw = t - kh * 3;
      tap_off[t] = S == 1 ? kh * p.pitch + kw
                          : ((kh & 1) * 2 + (kw & 1)) * p.plane_rows + (kh >> 1) * p.pitch + (kw >> 1);
    }
    // this lane's slots of a weight tile: row n = lane & 15, groups 2*(lane >> 4) and 2*(lane >> 4) + 1
    const int wn = lane & 15, wg0 = (lane >> 4) * 2;
    const uint32_t wslot0 = static_cast<uint32_t>(wn * 128 + ((((wg0 * 2) + (wn >> 3)) ^ (wn & 7)) << 4) + (wn & 7) * 2);
    const uint32_t wslot1 = static_cast<uint32_t>(wn * 128 + ((((wg0 * 2 + 2) + (wn >> 3)) ^ (wn & 7)) << 4) + (wn & 7) * 2);
    if (lane == 0) {
      wait_backoff(&ctrl->w_full, 0);
      tc_fence_after();
    }
    int c = 0, it = 0;        // chunk / tile of iteration g's depthwise part
    int pc = 0, pit = 0;      // chunk / tile of its project part (chunk g-1)
    for (int g = 0; g <= G; ++g) {
      if (g < G) {
        const int vc = min(64, p.Cexp - c * 64);
        const int wb = g % p.WB, eb = g % p.EB, d3b = g % p.D3B;
        uint8_t* wbuf = s_wb + wb * kWbBufBytes;
        if (lane == 0) wait_backoff(&ctrl->wb_empty[wb], ((static_cast<uint32_t>(g / p.WB)) & 1u) ^ 1u);
        __syncwarp();
        const __half* wsrc = s_dwh + c * 64 + wg0 * 16 + wn;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          *reinterpret_cast<__half*>(wbuf + t * kWbTileBytes + wslot0) = wsrc[t * CE];
          *reinterpret_cast<__half*>(wbuf + t * kWbTileBytes + wslot1) = wsrc[t * CE + 16];
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          wait_backoff(&ctrl->e_full[eb], static_cast<uint32_t>(g / p.EB) & 1u);
          wait_backoff(&ctrl->d3_empty[d3b], ((static_cast<uint32_t>(g / p.D3B)) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t ea = smem_u32(s_e + eb * e_buf_bytes);
          const uint32_t wa = smem_u32(wbuf);
          const uint32_t d3 = tmem_base + static_cast<uint32_t>(d3_col + d3b * 64);
          const int ngroups = vc >> 4;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint64_t da = make_smem_desc_sw128(ea + static_cast<uint32_t>(tap_off[t]) * 128u);
            const uint64_t db = make_smem_desc_sw128(wa + static_cast<uint32_t>(t) * kWbTileBytes);
            for (int q = 0; q < ngroups; ++q)
              umma_f16_ss(d3 + static_cast<uint32_t>(q * 16), da + static_cast<uint64_t>(q * 2),
                          db + static_cast<uint64_t>(q * 2), idesc_dw, t != 0 ? 1u : 0u);
          }
          umma_commit(&ctrl->e_empty[eb]);
          umma_commit(&ctrl->wb_empty[wb]);
          umma_commit(&ctrl->d3_full[d3b]);
        }
        if (++c == p.nc) {
          c = 0;
          ++it;
        }
      }
      if (g >= 1 && lane == 0) {
        const int gp = g - 1;
        const int ab = gp % p.AB, d2b = pit % p.D2B;
        if (pc == 0) wait_backoff(&ctrl->d2_empty[d2b], ((static_cast<uint32_t>(pit / p.D2B)) & 1u) ^ 1u);
        wait_backoff(&ctrl->a2_full[ab], static_cast<uint32_t>(gp / p.AB) & 1u);
        tc_fence_after();
        const int vc = min(64, p.Cexp - pc * 64);
        const int ks = (vc + 15) >> 4;
        const uint64_t da = make_smem_desc_sw128(smem_u32(s_a2 + ab * 16384));
        const uint64_t db = make_smem_desc_sw128(smem_u32(s_w2 + pc * w2_chunk));
        const uint32_t d2 = tmem_base + static_cast<uint32_t>(d2_col + d2b * 64);
        for (int k = 0; k < ks; ++k)
          umma_f16_ss(d2, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc2,
                      (pc | k) != 0 ? 1u : 0u);
        umma_commit(&ctrl->a2_empty[ab]);
        if (pc == p.nc - 1) umma_commit(&ctrl->d2_full[d2b]);
      }
      if (g >= 1) {
        if (++pc == p.nc) {
          pc = 0;
          ++pit;
        }
      }
    }
  } else {
    const int idx = (warp - kFirstEpiWarp) >> 2;   // 0..3: which of the four warps sharing this TMEM lane quarter
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    if (idx < nA) {
      // ============================ expand epilogue: D1 -> E (one swizzled 128-byte row per halo pixel) ============================
      int er[3], by[3], bx[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const int row = m * 128 + quarter * 32 + lane;
        by[m] = row / p.BW;
        bx[m] = row - by[m] * p.BW;
        er[m] = S == 1 ? row : ((by[m] & 1) * 2 + (bx[m] & 1)) * p.plane_rows + (by[m] >> 1) * p.pitch + (bx[m] >> 1);
      }
      int g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img;
        const int rr = tile - n * tiles_per_img;
        const int th_i = rr / p.tiles_w, tw_i = rr - th_i * p.tiles_w;
        const int iy0 = th_i * p.TH * S - 1, ix0 = tw_i * p.TW * S - 1;
        bool pvalid[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const int iy = iy0 + by[m], ix = ix0 + bx[m];
          pvalid[m] = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        }
        for (int c = 0; c < p.nc; ++c, ++g) {
          const int vc = min(64, p.Cexp - c * 64);
          const int eb = g % p.EB;
          mbar_wait(&ctrl->d1_full[g & 1], (static_cast<uint32_t>(g) >> 1) & 1u);
          tc_fence_after();
          mbar_wait(&ctrl->e_empty[eb], (static_cast<uint32_t>(g / p.EB) & 1u) ^ 1u);
          uint8_t* e_buf = s_e + eb * e_buf_bytes;
#pragma unroll
          for (int m = 0; m < 3; ++m) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              const int u = m * 2 + ch;
              if (m < p.Mtiles && (u % nA) == idx && ch * 32 < vc && m * 128 + quarter * 32 < p.n_rows) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + lane_sel + static_cast<uint32_t>((g & 1) * d1_cols + m * 64 + ch * 32), v);
                tmem_ld_wait();
                if (m * 128 + quarter * 32 + lane < p.n_rows)
                  bias_relu6_store(v, s_b1 + c * 64 + ch * 32, e_buf + er[m] * 128, er[m], ch * 4, !pvalid[m]);
              }
            }
          }
          tc_fence_before();
          fence_proxy_async();   // E is read by the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&ctrl->d1_empty[g & 1]);
            mbar_arrive(&ctrl->e_full[eb]);
          }
        }
      }
    } else {
      // ============================ depthwise epilogue (D3 -> A2) + project epilogue (D2 -> output) ============================
      const int idxB = idx - nA;
      const int bt = idxB * 128 + quarter * 32 + lane;   // thread index inside the depthwise-side group (only bt == 0 matters)
      const int m_row = quarter * 32 + lane;         // accumulator row of this thread: m = oy * pitch + ox
      const int oy_l = m_row / p.pitch, ox_l = m_row - oy_l * p.pitch;
      const bool m_ok = oy_l < p.TH && ox_l < p.TW;
      const bool q_active = quarter * 32 < p.m_max;
      const int srow_i = oy_l * p.TW + ox_l;         // row of the output staging tile (TMA box order)
      pdl_wait_prior_grid();

      auto project_epilogue = [&](int pit, int tile) {
        const int n = tile / tiles_per_img;
        const int rr = tile - n * tiles_per_img;
        const int th_i = rr / p.tiles_w, tw_i = rr - th_i * p.tiles_w;
        const int d2b = pit % p.D2B;
        mbar_wait(&ctrl->d2_full[d2b], static_cast<uint32_t>(pit / p.D2B) & 1u);
        tc_fence_after();
        if (bt == 0) tma_store_wait_read0();    // the previous tile's store has released the staging tile
        named_barrier_sync(kProjBarrier, nB * 128);
        if (q_active) {
          const int oy = th_i * p.TH + oy_l, ox = tw_i * p.TW + ox_l;
          const bool valid = m_ok && oy < p.Ho && ox < p.Wo;
          const __half* rp = nullptr;
          if (p.residual != nullptr && valid)
            rp = p.residual + ((static_cast<long long>(n) * p.Ho + oy) * p.Wo + ox) * p.res_stride;
          uint8_t* srow = s_out + srow_i * 128;
          const int ncg = p.cout_pad >> 4;
          for (int cg = idxB; cg < ncg; cg += nB) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_base + lane_sel + static_cast<uint32_t>(d2_col + d2b * 64 + cg * 16), v);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int co = cg * 16 + h * 8;
                const float4 ba = *reinterpret_cast<const float4*>(s_b3 + co);
                const float4 bb = *reinterpret_cast<const float4*>(s_b3 + co + 4);
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
                *reinterpret_cast<uint4*>(srow + (((cg * 2 + h) ^ (srow_i & 7)) << 4)) = ov;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->d2_empty[d2b]);
        fence_proxy_async();
        named_barrier_sync(kProjBarrier, nB * 128);
        if (bt == 0) {
          tma_store_4d(&maps.out, s_out, 0, tw_i * p.TW, th_i * p.TH, n);
          tma_store_commit();
        }
      };

      int it = 0, g = 0, prev_tile = -1;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        for (int c = 0; c < p.nc; ++c, ++g) {
          const int vc = min(64, p.Cexp - c * 64);
          const int d3b = g % p.D3B, ab = g % p.AB;
          mbar_wait(&ctrl->d3_full[d3b], static_cast<uint32_t>(g / p.D3B) & 1u);
          tc_fence_after();
          mbar_wait(&ctrl->a2_empty[ab], (static_cast<uint32_t>(g / p.AB) & 1u) ^ 1u);
          if (q_active) {
            uint8_t* arow = s_a2 + ab * 16384 + m_row * 128;
            for (int ch = idxB; ch < 2; ch += nB) {
              if (ch * 32 < vc) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + lane_sel + static_cast<uint32_t>(d3_col + d3b * 64 + ch * 32), v);
                tmem_ld_wait();
                bias_relu6_store(v, s_b2 + c * 64 + ch * 32, arow, m_row, ch * 4, false);
              }
            }
          }
          tc_fence_before();
          fence_proxy_async();   // A2 is read by the tensor core
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&ctrl->d3_empty[d3b]);
            mbar_arrive(&ctrl->a2_full[ab]);
          }
          // the previous tile's project epilogue runs one chunk late, so that its accumulator is complete by then
          if (c == 0 && prev_tile >= 0) project_epilogue(it - 1, prev_tile);
        }
        prev_tile = tile;
      }
      if (prev_tile >= 0) project_epilogue(it - 1, prev_tile);
      if (bt == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int round_up(int a, int b) { return (a + b - 1) / b * b; }

// shared-memory layout for a given tile + buffer configuration; returns total bytes
int mb3_layout(Mb3Params* p) {
  int off = p->XB * p->Mtiles * 16384;
  p->off_w1 = off;
  off += p->nc * 8192;
  p->off_w2 = off;
  off += p->nc * p->cout_pad * 128;
  off = round_up(off, 1024);
  p->off_a2 = off;
  off += p->AB * 16384;
  p->off_out = off;
  off += 16384;
  p->off_e = off;
  off += p->EB * p->e_rows * 128;
  p->off_wb = off;
  off += p->WB * kWbBufBytes;
  p->off_c = off;
  off += p->nc * 64 * (18 + 4 + 4) + 64 * 4;
  off = round_up(off, 16);
  p->off_ctrl = off;
  off += 256;
  p->smem = off;
  return off;
}

}  // namespace

bool mbconv3_plan(Mb3Params* p) {
  if (p->S != 1 && p->S != 2) return false;
  if (p->Cin < 8 || p->Cin > 64 || p->Cin % 8 != 0) return false;
  if (p->Cout < 8 || p->Cout > 64 || p->Cout % 8 != 0) return false;
  if (p->Cexp < 16 || p->Cexp % 16 != 0) return false;
  if (p->H < 4 || p->W < 4 || p->N < 1) return false;
  const int S = p->S;
  p->Ho = (p->H - 1) / S + 1;
  p->Wo = (p->W - 1) / S + 1;
  p->nc = (p->Cexp + 63) / 64;
  p->k1steps = (p->Cin + 15) / 16;
  p->cout_pad = (p->Cout + 15) / 16 * 16;
  int force_tw = 0, force_th = 0;
  if (const char* e = getenv("AF_MB3_TILE")) sscanf(e, "%d,%d", &force_tw, &force_th);
  // buffer configurations in order of preference: {XB, EB, AB, WB} and the cost penalty of each
  static const int cfgs[6][4] = {{2, 2, 2, 2}, {1, 2, 2, 2}, {1, 2, 2, 1}, {1, 2, 1, 1}, {1, 1, 2, 1}, {1, 1, 1, 1}};
  static const double pen[6] = {1.0, 1.04, 1.07, 1.15, 1.3, 1.4};
  double best = -1.0;
  Mb3Params bestp = *p;
  for (int th = 1; th <= 32; ++th) {
    for (int tw = 2; tw <= 64; ++tw) {
      if (force_tw > 0 && (tw != force_tw || th != force_th)) continue;
      Mb3Params q = *p;
      q.TW = tw;
      q.TH = th;
      q.pitch = S == 1 ? tw + 2 : tw + 1;
      q.m_max = (th - 1) * q.pitch + tw;
      if (q.m_max > 128) continue;
      q.BW = S == 1 ? tw + 2 : 2 * tw + 1;
      q.BH = S == 1 ? th + 2 : 2 * th + 1;
      q.n_rows = q.BW * q.BH;
      q.Mtiles = (q.n_rows + 127) / 128;
      if (q.Mtiles > 3 || q.BW > 256 || q.BH > 256) continue;
      if (S == 1) {
        q.plane_rows = 0;
        q.e_rows = round_up(q.n_rows > 130 + 2 * q.BW ? q.n_rows : 130 + 2 * q.BW, 8);
      } else {
        int pr = (th + 1) * q.pitch;
        while ((pr & 7) != 4) ++pr;
        q.plane_rows = pr;
        const int last = pr > 130 + q.pitch ? pr : 130 + q.pitch;
        q.e_rows = round_up(3 * pr + last, 8);
      }
      q.tiles_w = (q.Wo + tw - 1) / tw;
      q.tiles_h = (q.Ho + th - 1) / th;
      q.D3B = q.Mtiles <= 2 ? 2 : 1;
      q.D2B = q.Mtiles <= 2 ? 2 : 1;
      q.nA = q.Mtiles <= 2 ? 2 : 3;
      const double tiles = static_cast<double>(q.tiles_w) * q.tiles_h;
      // per-tile work: expand epilogue rows (both lane-quarter rounds), depthwise / project epilogue rows, fixed cost
      const double work = tiles * (q.n_rows * 1.0 / q.nA * 2 + round_up(q.m_max, 32) * 1.0 / (4 - q.nA) * 2 + 48);
      for (int k = 0; k < 6; ++k) {
        q.XB = cfgs[k][0];
        q.EB = cfgs[k][1];
        q.AB = cfgs[k][2];
        q.WB = cfgs[k][3];
        if (mb3_layout(&q) > 227 * 1024) continue;
        const double cost = work * pen[k];
        if (best < 0 || cost < best) {
          best = cost;
          bestp = q;
        }
        break;
      }
    }
  }
  if (best < 0) return false;
  *p = bestp;
  return true;
}

cudaError_t launch_mbconv3(const MbTensorMaps& maps, const Mb3Params& p, int sm_count, cudaStream_t stream) {
  static_assert(sizeof(Mb3Ctrl) <= 256, "ctrl block too large");
  using Kern = void (*)(const MbTensorMaps, const Mb3Params);
  Kern kern = p.S == 1 ? mbconv3_kernel<1> : mbconv3_kernel<2>;
  static bool attr_set[64][2] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_set[dev][p.S - 1]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev][p.S - 1] = true;
  }
  const long long n_tiles = 1LL * p.tiles_w * p.tiles_h * p.N;
  int grid = n_tiles < sm_count ? static_cast<int>(n_tiles) : sm_count;
  if (grid < 1) grid = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kMb3Threads);
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
