// Fused inverted-residual block kernel (see mbconv_fused.cuh for the data flow and the reference call sites).
#include "mbconv_fused.cuh"

#include <cstdlib>

#include "dw_strip.cuh"
#include "ptx.cuh"

namespace af {

using namespace ptx;

namespace {

constexpr int kEPitch = 144;      // bytes per pixel of the expanded tile E: 64 fp16 + 16 B pad (conflict-free 16-B row writes)
constexpr int kD2Col = 384;       // TMEM column of the project accumulator (D1 buffers: 2 x Mtiles x 64 <= 384 columns)
constexpr int kGroupThreads = 256;   // threads of each compute group (8 warps)
constexpr int kGroupWarps = kGroupThreads / 32;
constexpr int kFirstGroupWarp = 3;   // warps 0-2: TMA producer, expand MMA issuer, project MMA issuer
static_assert(kMbThreads == (kFirstGroupWarp + 2 * kGroupWarps) * 32, "thread roles");
constexpr int kEpilogueBarrier = 1;  // named barrier of the epilogue group (staging tile hand-over)

struct __align__(8) MbCtrl {
  uint64_t x_full[2], x_empty[2];
  uint64_t w_full;
  uint64_t d1_full[2], d1_empty[2];
  uint64_t e_full[2], e_empty[2];
  uint64_t a2_full[2], a2_empty[2];
  uint64_t d2_full[2], d2_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// tile index -> (tile column, tile row, image) as a mixed-radix counter advanced by gridDim.x per persistent-loop
// iteration (no per-tile integer divisions, cf. TileCursor in conv_gemm.cu)
struct MbCursor {
  int tw, th, n, d_tw, d_th, d_n;
  __device__ __forceinline__ void init(int tile, int step, const MbParams& p) {
    tw = tile % p.tiles_w;
    int r = tile / p.tiles_w;
    th = r % p.tiles_h;
    n = r / p.tiles_h;
    d_tw = step % p.tiles_w;
    r = step / p.tiles_w;
    d_th = r % p.tiles_h;
    d_n = r / p.tiles_h;
  }
  __device__ __forceinline__ void advance(const MbParams& p) {
    tw += d_tw;
    int c = tw >= p.tiles_w ? 1 : 0;
    tw -= c ? p.tiles_w : 0;
    th += d_th + c;
    c = th >= p.tiles_h ? 1 : 0;
    th -= c ? p.tiles_h : 0;
    n += d_n + c;
  }
};

__device__ __forceinline__ void wait_backoff(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(32);
}
#ifndef AF_MB_EPI_SLEEP
#define AF_MB_EPI_SLEEP 0
#endif
#ifndef AF_MB_DW_SLEEP
#define AF_MB_DW_SLEEP 0
#endif
template <int NS>
__device__ __forceinline__ void group_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
    if (NS > 0) __nanosleep(NS);
  }
}

template <int S>
__global__ void __launch_bounds__(kMbThreads, 1)
mbconv_fused_kernel(const __grid_constant__ MbTensorMaps maps, const MbParams p) {
  constexpr int RO = S == 1 ? 4 : 2;   // output rows per depthwise strip (256 strips of work per 64-channel chunk)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* s_x = smem;
  uint8_t* s_w1 = smem + p.off_w1;
  uint8_t* s_w2 = smem + p.off_w2;
  uint8_t* s_a2 = smem + p.off_a2;
  uint8_t* s_out = smem + p.off_out;
  uint8_t* s_e = smem + p.off_e;
  float* s_f = reinterpret_cast<float*>(smem + p.off_f32);
  MbCtrl* ctrl = reinterpret_cast<MbCtrl*>(smem + p.off_ctrl);
  const int CE = p.nc * 64;
  float* s_dw = s_f;              // [9][CE]
  float* s_b1 = s_f + 9 * CE;     // [CE]
  float* s_b2 = s_b1 + CE;        // [CE]
  float* s_b3 = s_b2 + CE;        // [64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x_buf_bytes = p.Mtiles * 16384;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int n_tiles = tiles_per_img * p.N;
  const uint32_t w2_chunk = static_cast<uint32_t>(p.cout_pad) * 128u;
  const int d1_cols = p.Mtiles * 64;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctrl->x_full[i], 1);
      mbar_init(&ctrl->x_empty[i], 1);
      mbar_init(&ctrl->d1_full[i], 1);
      mbar_init(&ctrl->d1_empty[i], kGroupWarps);
      mbar_init(&ctrl->e_full[i], kGroupWarps);
      mbar_init(&ctrl->e_empty[i], kGroupWarps);
      mbar_init(&ctrl->a2_full[i], kGroupWarps);
      mbar_init(&ctrl->a2_empty[i], 1);
      mbar_init(&ctrl->d2_full[i], 1);
      mbar_init(&ctrl->d2_empty[i], kGroupWarps);
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
  if (warp >= kFirstGroupWarp) {
    // launch constants -> shared memory; the A2 operand buffers start as zeros so that channel columns a partial
    // chunk never writes hold finite values (their weights are zero)
    const int ct = threadIdx.x - kFirstGroupWarp * 32;
    for (int i = ct; i < 2 * 16384 / 16; i += 2 * kGroupThreads)
      reinterpret_cast<uint4*>(s_a2)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = ct; i < 9 * CE; i += 2 * kGroupThreads) s_dw[i] = p.dw_w[i];
    for (int i = ct; i < CE; i += 2 * kGroupThreads) {
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

  if (warp == 0) {
    // ============================ TMA producer: weights once, then one input window per tile ============================
    {   // converged warp: all lanes run the loop, one elected lane issues (ptx.cuh, *_elect)
      mbar_arrive_expect_tx_elect(&ctrl->w_full, static_cast<uint32_t>(p.nc) * (8192u + w2_chunk));
      for (int c = 0; c < p.nc; ++c) tma_load_2d_elect(s_w1 + c * 8192, &maps.w1, &ctrl->w_full, 0, c * 64);
      for (int c = 0; c < p.nc; ++c) tma_load_2d_elect(s_w2 + c * w2_chunk, &maps.w2, &ctrl->w_full, c * 64, 0);
      pdl_wait_prior_grid();
      int it = 0;
      MbCursor cur;
      cur.init(blockIdx.x, gridDim.x, p);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it, cur.advance(p)) {
        const int xb = it & (p.XB - 1);   // XB is 1 or 2
        const uint32_t ph = static_cast<uint32_t>(it >> (p.XB - 1)) & 1u;
        while (!mbar_try_wait(&ctrl->x_empty[xb], ph ^ 1u)) __nanosleep(256);
        const int n = cur.n, th_i = cur.th, tw_i = cur.tw;
        mbar_arrive_expect_tx_elect(&ctrl->x_full[xb], static_cast<uint32_t>(p.n_rows) * 128u);
        tma_load_4d_elect(s_x + xb * x_buf_bytes, &maps.x, &ctrl->x_full[xb], 0, tw_i * p.TW * S - 1, th_i * p.TH * S - 1, n);
      }
    }
  } else if (warp == 1) {
    // ============================ expand MMA issuer: D1[g & 1] = X * W1[chunk]^T as soon as the buffer is free ============================
    {   // converged warp: all lanes run the loop, one elected lane issues (ptx.cuh, *_elect)
      const uint32_t idesc1 = make_idesc_f16_f32(128, 64);
      const int my_tiles = static_cast<int>(blockIdx.x) < n_tiles
                               ? (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                     static_cast<int>(gridDim.x)
                               : 0;
      const int G = my_tiles * p.nc;   // channel chunks this CTA goes through, across all its tiles
      wait_backoff(&ctrl->w_full, 0);
      tc_fence_after();
      // bias-in-MMA: lane l owns halo rows l, l+32, ..; their (by, bx) inside the halo box never change
      int one_by[12], one_bx[12];
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const int r = lane + 32 * j;
        one_by[j] = r / p.BW;
        one_bx[j] = r - one_by[j] * p.BW;
      }
      int it = 0, c = 0;
      MbCursor cur;
      cur.init(blockIdx.x, gridDim.x, p);
      for (int g = 0; g < G; ++g) {
        const int xb = it & (p.XB - 1);   // XB is 1 or 2
        if (c == 0) {
          wait_backoff(&ctrl->x_full[xb], static_cast<uint32_t>(it >> (p.XB - 1)) & 1u);
          if (p.bias_col >= 0) {
            // plant the constant-1 channel pair in every in-image pixel of the freshly landed window (TMA zero-filled
            // the channels past Cin and the pixels outside the image): X * [W1 | bias_hi | bias_lo]^T then yields
            // x W1^T + bias inside the image and exactly 0 outside it
            const int iy0 = cur.th * p.TH * S - 1, ix0 = cur.tw * p.TW * S - 1;
            uint8_t* xt = s_x + xb * x_buf_bytes;
            const uint32_t chunk = static_cast<uint32_t>(p.bias_col >> 3), sub = static_cast<uint32_t>(p.bias_col & 7) * 2u;
#pragma unroll
            for (int j = 0; j < 12; ++j) {
              const int r = lane + 32 * j;
              const int iy = iy0 + one_by[j], ix = ix0 + one_bx[j];
              if (r < p.n_rows && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
                *reinterpret_cast<uint32_t*>(xt + r * 128 + ((chunk ^ (static_cast<uint32_t>(r) & 7u)) << 4) + sub) =
                    0x3C003C00u;   // fp16 (1.0, 1.0)
            }
            fence_proxy_async();   // generic-proxy writes -> visible to the tensor core
            __syncwarp();
          }
        }
        wait_backoff(&ctrl->d1_empty[g & 1], ((static_cast<uint32_t>(g) >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t la0 = smem_desc_lo(smem_u32(s_x + xb * x_buf_bytes));
        const uint32_t lb = smem_desc_lo(smem_u32(s_w1 + c * 8192));
        for (int m = 0; m < p.Mtiles; ++m) {
          const uint32_t la = la0 + static_cast<uint32_t>(m) * (16384u >> 4);
          const uint32_t d = tmem_base + static_cast<uint32_t>((g & 1) * d1_cols + m * 64);
          for (int k = 0; k < p.k1steps; ++k)
            umma_f16_ss_lo_elect(d, la + static_cast<uint32_t>(k * 2), lb + static_cast<uint32_t>(k * 2), idesc1,
                                 k != 0 ? 1u : 0u);
        }
        umma_commit_elect(&ctrl->d1_full[g & 1]);
        if (c == p.nc - 1) umma_commit_elect(&ctrl->x_empty[xb]);   // the tile's input window has been consumed
        if (++c == p.nc) {
          c = 0;
          ++it;
          cur.advance(p);
        }
      }
    }
  } else if (warp == 2) {
    // ============================ project MMA issuer: D2[it & 1] += A2[g & 1] * W2[:, chunk]^T ============================
    {   // converged warp: all lanes run the loop, one elected lane issues (ptx.cuh, *_elect)
      const uint32_t idesc2 = make_idesc_f16_f32(128, static_cast<uint32_t>(p.cout_pad));
      const int my_tiles = static_cast<int>(blockIdx.x) < n_tiles
                               ? (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                     static_cast<int>(gridDim.x)
                               : 0;
      const int G = my_tiles * p.nc;
      wait_backoff(&ctrl->w_full, 0);
      tc_fence_after();
      int it = 0, c = 0;
      for (int g = 0; g < G; ++g) {
        if (c == 0) {
          // the accumulator buffer of this tile was last read by the epilogue of tile it-2
          wait_backoff(&ctrl->d2_empty[it & 1], ((static_cast<uint32_t>(it) >> 1) & 1u) ^ 1u);
        }
        wait_backoff(&ctrl->a2_full[g & 1], (static_cast<uint32_t>(g) >> 1) & 1u);
        tc_fence_after();
        const int vc = min(64, p.Cexp - c * 64);
        const int ks = (vc + 15) >> 4;
        const uint32_t la = smem_desc_lo(smem_u32(s_a2 + (g & 1) * 16384));
        const uint32_t lb = smem_desc_lo(smem_u32(s_w2 + c * w2_chunk));
        const uint32_t d2 = tmem_base + static_cast<uint32_t>(kD2Col + (it & 1) * 64);
        for (int k = 0; k < ks; ++k)
          umma_f16_ss_lo_elect(d2, la + static_cast<uint32_t>(k * 2), lb + static_cast<uint32_t>(k * 2), idesc2,
                               (c | k) != 0 ? 1u : 0u);
        umma_commit_elect(&ctrl->a2_empty[g & 1]);
        if (c == p.nc - 1) umma_commit_elect(&ctrl->d2_full[it & 1]);
        if (++c == p.nc) {
          c = 0;
          ++it;
        }
      }
    }
  } else if (warp < kFirstGroupWarp + kGroupWarps) {
    // ============================ epilogue group (8 warps): D1 -> E per chunk, D2 -> output per tile ============================
    const int et = threadIdx.x - kFirstGroupWarp * 32;   // 0..255
    const int quarter = warp & 3;             // TMEM lane quarter this warp may access
    const int colhalf = (warp - kFirstGroupWarp) >> 2;   // which 32 of a chunk's 64 columns (expand) / which 16-column groups (project)
    int by[3], bx[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      const int row = m * 128 + quarter * 32 + lane;
      by[m] = row / p.BW;
      bx[m] = row - by[m] * p.BW;
    }
    const int prow = quarter * 32 + lane;     // output pixel (row of the project accumulator) of this thread
    const int pth = prow / p.TW, ptw = prow - pth * p.TW;
    const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
    pdl_wait_prior_grid();

    // project epilogue of tile `pit` (this CTA's pit-th tile): D2 (TMEM) -> +bias (+ residual) -> fp16 -> staging -> TMA store
    auto project_epilogue = [&](int pit, int n, int th_i, int tw_i) {
      group_wait<AF_MB_EPI_SLEEP>(&ctrl->d2_full[pit & 1], (static_cast<uint32_t>(pit) >> 1) & 1u);
      tc_fence_after();
      if (et == 0) tma_store_wait_read0();    // the previous tile's store has released the staging tile
      named_barrier_sync(kEpilogueBarrier, kGroupThreads);
      if (quarter * 32 < p.TW * p.TH) {
        const int oy = th_i * p.TH + pth, ox = tw_i * p.TW + ptw;
        const bool valid = prow < p.TW * p.TH && oy < p.Ho && ox < p.Wo;
        const __half* rp = nullptr;
        if (p.residual != nullptr && valid)
          rp = p.residual + ((static_cast<long long>(n) * p.Ho + oy) * p.Wo + ox) * p.res_stride;
        uint8_t* srow = s_out + prow * 128;
        const int ncg = p.cout_pad >> 4;
        for (int cg = colhalf; cg < ncg; cg += 2) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(tmem_base + lane_sel + static_cast<uint32_t>(kD2Col + (pit & 1) * 64 + cg * 16), v);
          tmem_ld_wait();
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
            *reinterpret_cast<uint4*>(srow + (((cg * 2 + h) ^ (prow & 7)) << 4)) = ov;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctrl->d2_empty[pit & 1]);
      fence_proxy_async();
      named_barrier_sync(kEpilogueBarrier, kGroupThreads);
      if (et == 0) {
        tma_store_4d(&maps.out, s_out, 0, tw_i * p.TW, th_i * p.TH, n);
        tma_store_commit();
      }
    };

    int it = 0, g = 0, prev_tile = -1, prev_n = 0, prev_th = 0, prev_tw = 0;
    MbCursor cur;
    cur.init(blockIdx.x, gridDim.x, p);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it, cur.advance(p)) {
      const int n = cur.n, th_i = cur.th, tw_i = cur.tw;
      const int iy0 = th_i * p.TH * S - 1, ix0 = tw_i * p.TW * S - 1;
      bool pvalid[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const int iy = iy0 + by[m], ix = ix0 + bx[m];
        pvalid[m] = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
      }
      for (int c = 0; c < p.nc; ++c, ++g) {
        const int vc = min(64, p.Cexp - c * 64);
        const int eb = g & (p.EB - 1);   // EB is 1 or 2
        // ---- expand epilogue: D1 (TMEM) -> +bias, ReLU6, zero outside the image -> E (smem, fp16)
        group_wait<AF_MB_EPI_SLEEP>(&ctrl->d1_full[g & 1], (static_cast<uint32_t>(g) >> 1) & 1u);
        tc_fence_after();
        group_wait<AF_MB_EPI_SLEEP>(&ctrl->e_empty[eb], (static_cast<uint32_t>(g >> (p.EB - 1)) & 1u) ^ 1u);
        if (colhalf * 32 < vc) {
          uint8_t* e_buf = s_e + eb * p.e_bytes;
#pragma unroll
          for (int m = 0; m < 3; ++m) {
            if (m < p.Mtiles && m * 128 + quarter * 32 < p.n_rows) {
              uint32_t v[32];
              tmem_ld_32x32b_x32(tmem_base + lane_sel + static_cast<uint32_t>((g & 1) * d1_cols + m * 64 + colhalf * 32), v);
              tmem_ld_wait();
              const int row = m * 128 + quarter * 32 + lane;
              if (row < p.n_rows) {
                uint8_t* erow = e_buf + row * kEPitch + colhalf * 64;
                const float* b1 = s_b1 + c * 64 + colhalf * 32;
                if (p.bias_col >= 0) {
                  // bias and the zero padding already came out of the MMA: ReLU6 + convert only
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    uint4 ov;
                    __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                      oh2[j] = __hmin2(floats2half2_relu(__uint_as_float(v[i * 8 + 2 * j]), __uint_as_float(v[i * 8 + 2 * j + 1])),
                                       __float2half2_rn(6.f));
                    *reinterpret_cast<uint4*>(erow + i * 16) = ov;
                  }
                } else {
                  const bool ok = pvalid[m];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float4 ba = *reinterpret_cast<const float4*>(b1 + i * 8);
                    const float4 bb = *reinterpret_cast<const float4*>(b1 + i * 8 + 4);
                    const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                    uint4 ov;
                    __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      const __half2 h = floats2half2_relu(__uint_as_float(v[i * 8 + 2 * j]) + bv[2 * j],
                                                          __uint_as_float(v[i * 8 + 2 * j + 1]) + bv[2 * j + 1]);
                      oh2[j] = __hmin2(h, __float2half2_rn(6.f));
                    }
                    if (!ok) ov = make_uint4(0u, 0u, 0u, 0u);   // depthwise zero padding lives in the expanded domain
                    *reinterpret_cast<uint4*>(erow + i * 16) = ov;
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&ctrl->d1_empty[g & 1]);
          mbar_arrive(&ctrl->e_full[eb]);
        }
        // the previous tile's project epilogue runs one chunk late, so that its accumulator is complete by then
        if (c == 0 && prev_tile >= 0) project_epilogue(it - 1, prev_n, prev_th, prev_tw);
      }
      prev_tile = tile;
      prev_n = n;
      prev_th = th_i;
      prev_tw = tw_i;
    }
    if (prev_tile >= 0) project_epilogue(it - 1, prev_n, prev_th, prev_tw);
    if (et == 0) tma_store_wait_all();
  } else {
    // ============================ depthwise group (8 warps): E -> depthwise 3x3 + bias + ReLU6 -> A2 ============================
    const int dt = threadIdx.x - kFirstGroupWarp * 32 - kGroupThreads;   // 0..255
    const int pairs = p.TW >> 1;                 // TW is 8, 16 or 32: powers of two, so the index splits are shifts
    const int pairs_log2 = 31 - __clz(pairs);
    const int q_count = pairs * p.strips;
    const int my_tiles = static_cast<int>(blockIdx.x) < n_tiles
                             ? (n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                   static_cast<int>(gridDim.x)
                             : 0;
    const int G = my_tiles * p.nc;
    int c = 0;
    for (int g = 0; g < G; ++g) {
      const int vc = min(64, p.Cexp - c * 64);
      const int eb = g & (p.EB - 1);   // EB is 1 or 2
      const int n_ch = vc >> 2;             // 4-channel groups of this chunk: 16, 8 or 4
      const int n_ch_log2 = 31 - __clz(n_ch);
      const int ch4 = dt & (n_ch - 1);
      const int q0 = dt >> n_ch_log2;
      const int q_step = kGroupThreads >> n_ch_log2;
      const int ce = c * 64 + ch4 * 4;
      float2 w[9][2], bias[2];
      if (q0 < q_count) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 wv = *reinterpret_cast<const float4*>(s_dw + t * CE + ce);
          w[t][0] = make_float2(wv.x, wv.y);
          w[t][1] = make_float2(wv.z, wv.w);
        }
        const float4 b2 = *reinterpret_cast<const float4*>(s_b2 + ce);
        bias[0] = make_float2(b2.x, b2.y);
        bias[1] = make_float2(b2.z, b2.w);
      }
      group_wait<AF_MB_DW_SLEEP>(&ctrl->e_full[eb], static_cast<uint32_t>(g >> (p.EB - 1)) & 1u);
      group_wait<AF_MB_DW_SLEEP>(&ctrl->a2_empty[g & 1], ((static_cast<uint32_t>(g) >> 1) & 1u) ^ 1u);
      if (q0 < q_count) {
        const uint8_t* e_buf = s_e + eb * p.e_bytes;
        uint8_t* a2 = s_a2 + (g & 1) * 16384;
        for (int q = q0; q < q_count; q += q_step) {
          const int strip = q >> pairs_log2, xp = q & (pairs - 1);
          const uint8_t* in = e_buf + ((strip * RO * S) * p.BW + xp * 2 * S) * kEPitch + ch4 * 8;
          const int prow0 = strip * RO * p.TW + 2 * xp;
          uint8_t* out0 = a2 + prow0 * 128 + (((ch4 >> 1) ^ (prow0 & 7)) << 4) + (ch4 & 1) * 8;
          uint8_t* out1 = a2 + (prow0 + 1) * 128 + (((ch4 >> 1) ^ ((prow0 + 1) & 7)) << 4) + (ch4 & 1) * 8;
          dw_strip<S, RO>(in, kEPitch, p.BW * kEPitch, out0, out1, p.TW * 128, w, bias, 2);
        }
      }
      fence_proxy_async();                  // A2 writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ctrl->e_empty[eb]);
        mbar_arrive(&ctrl->a2_full[g & 1]);
      }
      if (++c == p.nc) c = 0;
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

bool mbconv_plan(MbParams* p) {
  if (p->S != 1 && p->S != 2) return false;
  if (p->Cin < 8 || p->Cin > 64 || p->Cin % 8 != 0) return false;
  if (p->Cout < 8 || p->Cout > 64 || p->Cout % 8 != 0) return false;
  if (p->Cexp < 16 || p->Cexp % 16 != 0) return false;
  const int tail = p->Cexp % 64;
  if (tail == 48) return false;               // chunk widths are 16, 32 or 64 channels
  if (p->H < 4 || p->W < 4 || p->N < 1) return false;
  p->Ho = (p->H - 1) / p->S + 1;
  p->Wo = (p->W - 1) / p->S + 1;
  p->nc = (p->Cexp + 63) / 64;
  p->k1steps = (p->Cin + 15) / 16;
  p->cout_pad = (p->Cout + 15) / 16 * 16;
  // output tile: 128 pixels (stride 1) or 64 pixels (stride 2: the halo'd input window must fit three 128-row MMA tiles)
  const int cand1[3][2] = {{16, 8}, {8, 16}, {32, 4}};
  const int cand2[2][2] = {{8, 8}, {16, 4}};
  long long best = -1;
  const int ncand = p->S == 1 ? 3 : 2;
  for (int i = 0; i < ncand; ++i) {
    const int tw = p->S == 1 ? cand1[i][0] : cand2[i][0];
    const int th = p->S == 1 ? cand1[i][1] : cand2[i][1];
    const long long tiles = 1LL * ((p->Wo + tw - 1) / tw) * ((p->Ho + th - 1) / th);
    if (best < 0 || tiles < best) {
      best = tiles;
      p->TW = tw;
      p->TH = th;
    }
  }
  p->strips = p->TH / (p->S == 1 ? 4 : 2);
  const int edge = p->S == 1 ? 2 : 1;
  p->BW = p->TW * p->S + edge;
  p->BH = p->TH * p->S + edge;
  p->n_rows = p->BW * p->BH;
  p->Mtiles = (p->n_rows + 127) / 128;
  if (p->Mtiles > 3) return false;
  p->tiles_w = (p->Wo + p->TW - 1) / p->TW;
  p->tiles_h = (p->Ho + p->TH - 1) / p->TH;
  p->e_bytes = (p->n_rows * kEPitch + 127) & ~127;
  // preferred: two E buffers (the epilogue group fills one while the depthwise group reads the other) and two input
  // windows; shrink to what fits in 227 KiB
  // when both double buffers do not fit, two input windows + one E buffer measured 1-2 % faster than the reverse on the
  // stride-2 blocks (b2 881 -> 864 us); AF_MB_EB_FIRST restores the round-1 order
  static const bool xb_first = getenv("AF_MB_EB_FIRST") == nullptr;
  const int try_xb[4] = {2, xb_first ? 2 : 1, xb_first ? 1 : 2, 1}, try_eb[4] = {2, xb_first ? 1 : 2, xb_first ? 2 : 1, 1};
  for (int t = 0; t < 4; ++t) {
    int off = try_xb[t] * p->Mtiles * 16384;
    p->off_w1 = off;
    off += p->nc * 8192;
    p->off_w2 = off;
    off += p->nc * p->cout_pad * 128;
    off = (off + 1023) & ~1023;
    p->off_a2 = off;
    off += 2 * 16384;
    p->off_out = off;
    off += 16384;
    p->off_e = off;
    off += try_eb[t] * p->e_bytes;
    p->off_f32 = off;
    off += (11 * p->nc * 64 + 64) * 4;
    off = (off + 15) & ~15;
    p->off_ctrl = off;
    off += 256;
    p->smem = off;
    p->XB = try_xb[t];
    p->EB = try_eb[t];
    if (off <= 227 * 1024) return true;
  }
  return false;
}

cudaError_t launch_mbconv_fused(const MbTensorMaps& maps, const MbParams& p, int sm_count, cudaStream_t stream) {
  static_assert(sizeof(MbCtrl) <= 256, "ctrl block too large");
  using Kern = void (*)(const MbTensorMaps, const MbParams);
  Kern kern = p.S == 1 ? mbconv_fused_kernel<1> : mbconv_fused_kernel<2>;
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
  cfg.blockDim = dim3(kMbThreads);
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
