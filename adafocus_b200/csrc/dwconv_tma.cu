// Depthwise 3x3 convolution (pad 1, stride 1 or 2) + folded BatchNorm + activation on NHWC fp16, TMA-staged (sm_100a).
// Replaces Conv2d(hidden, hidden, 3, stride, 1, groups=hidden) + BN + ReLU6 of ACT/models/mobilenet.py:56-59.
//
// HBM-bound layer, so the kernel is built around the memory system rather than the FMA pipes:
//  * persistent CTAs; every tile's input window {CB channels, TW*S+2, TH*S+2, NB images} arrives as ONE 4-D TMA box
//    (out-of-image halo = TMA zero fill = the conv padding; no bounds checks or address arithmetic in the hot loop),
//    double-buffered so the next tile's window is in flight while the current one is consumed;
//  * each thread owns 4 channels x 2 adjacent output columns x RO output rows and slides down the rows: an input row
//    is read once from shared memory (LDS.64), converted once, and feeds up to three output rows held in registers;
//    the 9x4 weights (BN scale folded in) live in registers as packed pairs and the math is FFMA2 (fma.rn.f32x2);
//  * results go to a staging tile in shared memory and leave with ONE TMA store per tile (partial tiles are clipped
//    by the hardware).
#include <cstdlib>

#include "dwconv_tma.cuh"
#include "dw_strip.cuh"
#include "ptx.cuh"

namespace af {

using namespace ptx;

namespace {

struct __align__(8) DwCtrl {
  uint64_t full[2];
};

template <int S, int RO>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_tma_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                     const DwTmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_in = smem;
  uint8_t* s_out = smem + 2 * p.in_bytes;
  DwCtrl* ctrl = reinterpret_cast<DwCtrl*>(s_out + 2 * p.out_bytes);

  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&ctrl->full[0], 1);
    mbar_init(&ctrl->full[1], 1);
    fence_mbar_init();
    tma_prefetch_desc(&in_map);
    tma_prefetch_desc(&out_map);
  }
  __syncthreads();

  // channel block of this CTA (fixed for its whole life: the grid is a multiple of n_cb) and the thread's 4 channels
  const int cb = blockIdx.x % p.n_cb;
  const int sp0 = blockIdx.x / p.n_cb;
  const int sp_step = gridDim.x / p.n_cb;
  const int n_chunks = p.CB >> 2;
  const int chunk = tid % n_chunks;
  const int q0 = tid / n_chunks;
  const int q_step = blockDim.x / n_chunks;
  const int c = cb * p.CB + chunk * 4;

  // weights and bias are launch constants (not produced by the previous kernel): load before the PDL wait
  float2 w[9][2], bias[2];
  {
    const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + c));
    const float4 bi = __ldg(reinterpret_cast<const float4*>(p.bias + c));
    bias[0] = make_float2(bi.x, bi.y);
    bias[1] = make_float2(bi.z, bi.w);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(p.w9c + t * p.C + c));
      w[t][0] = make_float2(wv.x * sc.x, wv.y * sc.y);
      w[t][1] = make_float2(wv.z * sc.z, wv.w * sc.w);
    }
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int n_sp = p.tiles_w * p.tiles_h * p.tiles_n;
  const int BW = p.TW * S + (S == 1 ? 2 : 1), BH = p.TH * S + (S == 1 ? 2 : 1);
  const int pix_b = p.CB * 2, row_b = BW * pix_b, img_b = BH * row_b;
  const int opix_b = p.CB * 2, orow_b = p.TW * opix_b, oimg_b = p.TH * orow_b;
  const int pairs = p.TW >> 1;
  const int q_count = pairs * p.strips * p.NB;
  // work item q = (nb * strips + strip) * pairs + xp, visited as q0, q0 + q_step, ..: a mixed-radix counter, so the
  // tile loop has no integer divisions
  const int xp0 = q0 % pairs, strip0 = (q0 / pairs) % p.strips, nb0 = q0 / (pairs * p.strips);
  const int dxp = q_step % pairs, dstrip = (q_step / pairs) % p.strips, dnb = q_step / (pairs * p.strips);

  auto issue_load = [&](int sp, int buf) {
    const int tw_i = sp % p.tiles_w;
    const int th_i = (sp / p.tiles_w) % p.tiles_h;
    const int tn_i = sp / (p.tiles_w * p.tiles_h);
    mbar_arrive_expect_tx(&ctrl->full[buf], static_cast<uint32_t>(p.NB * img_b));
    tma_load_4d(s_in + buf * p.in_bytes, &in_map, &ctrl->full[buf], cb * p.CB, tw_i * p.TW * S - 1,
                th_i * p.TH * S - 1, tn_i * p.NB);
  };

  if (tid == 0) {
    if (sp0 < n_sp) issue_load(sp0, 0);
    if (sp0 + sp_step < n_sp) issue_load(sp0 + sp_step, 1);
  }

  int it = 0;
  for (int sp = sp0; sp < n_sp; sp += sp_step, ++it) {
    const int buf = it & 1;
    mbar_wait(&ctrl->full[buf], (it >> 1) & 1);
    const uint8_t* tin = s_in + buf * p.in_bytes + chunk * 8;
    uint8_t* tout = s_out + buf * p.out_bytes + chunk * 8;
    int xp = xp0, strip = strip0, nb = nb0;
    for (int q = q0; q < q_count; q += q_step) {
      dw_strip<S, RO>(tin + nb * img_b + (strip * RO * S) * row_b + (xp * 2 * S) * pix_b, pix_b, row_b,
                      tout + nb * oimg_b + (strip * RO) * orow_b + (xp * 2) * opix_b,
                      tout + nb * oimg_b + (strip * RO) * orow_b + (xp * 2 + 1) * opix_b, orow_b, w, bias, p.act);
      xp += dxp;
      int carry = xp >= pairs ? 1 : 0;
      xp -= carry ? pairs : 0;
      strip += dstrip + carry;
      carry = strip >= p.strips ? 1 : 0;
      strip -= carry ? p.strips : 0;
      nb += dnb + carry;
    }
    fence_proxy_async();
    __syncthreads();   // every read of s_in[buf] and every write of s_out[buf] is done
    if (tid == 0) {
      const int tw_i = sp % p.tiles_w;
      const int th_i = (sp / p.tiles_w) % p.tiles_h;
      const int tn_i = sp / (p.tiles_w * p.tiles_h);
      tma_store_4d(&out_map, s_out + buf * p.out_bytes, cb * p.CB, tw_i * p.TW, th_i * p.TH, tn_i * p.NB);
      tma_store_commit();
      if (sp + 2 * sp_step < n_sp) issue_load(sp + 2 * sp_step, buf);
      tma_store_wait_read1();   // the previous tile's store has finished reading the other staging buffer
    }
    __syncthreads();
  }
  if (tid == 0) tma_store_wait_all();
}

}  // namespace

bool dwconv_tma_plan(int N, int H, int W, int C, int stride, DwTmaParams* out) {
  if (C % 8 != 0 || (stride != 1 && stride != 2) || N < 1 || H < 2 || W < 2) return false;
  int CB = 0;
  if (C % 64 == 0) CB = 64;
  else if (C % 48 == 0) CB = 48;
  else if (C % 32 == 0) CB = 32;
  else return false;
  const int threads = CB == 48 ? 192 : 256;
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const int edge = stride == 1 ? 2 : 1;
  const int tw_cand[] = {8, 14, 16, 28, 32};
  const int ro_cand[] = {4, 7};
  const int strips_cand[] = {1, 2, 4};
  const int nb_cand[] = {1, 2, 4, 8};
  double best = -1;
  DwTmaParams b = {};
  // first pass: tilings that leave room for two CTAs per SM; second pass: anything that fits
  for (int pass = 0; pass < 2 && best < 0; ++pass)
  for (int twc : tw_cand) {
    const int wo_even = (Wo + 1) & ~1;
    const int TW = twc < wo_even ? twc : wo_even;
    for (int RO : ro_cand) {
      for (int strips : strips_cand) {
        const int TH = RO * strips;
        if (TH - RO >= Ho) continue;                 // a whole strip would be outside the image
        for (int NB : nb_cand) {
          if (NB > 1 && (TH < Ho || TW < Wo)) continue;   // several images per tile only when one tile covers an image
          if (NB > N && NB > 1) continue;
          const int BW = TW * stride + edge, BH = TH * stride + edge;
          if (BW > 256 || BH > 256) continue;
          const int in_bytes = (NB * BH * BW * CB * 2 + 127) & ~127;
          const int out_bytes = (NB * TH * TW * CB * 2 + 127) & ~127;
          const int smem = 2 * in_bytes + 2 * out_bytes + 64;
          if (smem > (pass == 0 ? 110 : 200) * 1024) continue;
          const int items = (CB / 4) * (TW / 2) * strips * NB;
          const int passes = (items + threads - 1) / threads;
          const long long tiles = 1LL * ((Wo + TW - 1) / TW) * ((Ho + TH - 1) / TH) * ((N + NB - 1) / NB);
          double cost = static_cast<double>(tiles) * passes * ((RO - 1) * stride + 3 + 1);
          if (smem > 100 * 1024) cost *= 1.4;        // only one CTA per SM
          cost += tiles * 6.0;                       // per-tile fixed cost (barriers, TMA issue)
          if (best < 0 || cost < best) {
            best = cost;
            b.CB = CB; b.TW = TW; b.TH = TH; b.RO = RO; b.strips = strips; b.NB = NB;
            b.in_bytes = in_bytes; b.out_bytes = out_bytes; b.threads = threads; b.smem = smem;
            b.tiles_w = (Wo + TW - 1) / TW; b.tiles_h = (Ho + TH - 1) / TH; b.tiles_n = (N + NB - 1) / NB;
          }
        }
      }
    }
  }
  if (best < 0) return false;
  b.n_cb = C / CB;
  b.C = C;
  *out = b;
  return true;
}

cudaError_t launch_dwconv3x3_tma(const CUtensorMap& in_map, const CUtensorMap& out_map, const DwTmaParams& p,
                                 int stride, int sm_count, cudaStream_t stream) {
  using Kern = void (*)(const CUtensorMap, const CUtensorMap, const DwTmaParams);
  Kern kern = nullptr;
  if (stride == 1) kern = p.RO == 4 ? dwconv3x3_tma_kernel<1, 4> : dwconv3x3_tma_kernel<1, 7>;
  else kern = p.RO == 4 ? dwconv3x3_tma_kernel<2, 4> : dwconv3x3_tma_kernel<2, 7>;
  static bool attr_set[64][4] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const int ki = (stride - 1) * 2 + (p.RO == 4 ? 0 : 1);
  if (dev < 0 || dev >= 64 || !attr_set[dev][ki]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 64);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_set[dev][ki] = true;
  }
  const int ctas_per_sm = p.smem > 100 * 1024 ? 1 : 2;
  const long long n_sp = 1LL * p.tiles_w * p.tiles_h * p.tiles_n;
  long long grid = 1LL * sm_count * ctas_per_sm / p.n_cb * p.n_cb;
  if (grid < p.n_cb) grid = p.n_cb;
  if (grid > n_sp * p.n_cb) grid = n_sp * p.n_cb;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(p.threads);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = stream;
  cfg.attrs = nullptr;
  cfg.numAttrs = 0;
  return cudaLaunchKernelEx(&cfg, kern, in_map, out_map, p);
}

}  // namespace af
