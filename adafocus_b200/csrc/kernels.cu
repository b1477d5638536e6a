// HBM-bound helper kernels. All activations are NHWC fp16 unless the name says otherwise; math is fp32.
#include "kernels.cuh"

#include <cstdlib>

namespace af {

namespace {

constexpr int kThreads = 256;

// Launch with programmatic stream serialization: the kernel may be scheduled while the previous kernel of the stream
// drains; every kernel launched this way calls pdl_sync() before touching global memory.  kEarly = false launches
// normally: measured on B200, letting the large multi-wave kernels (depthwise conv, stem staging, pooling) start early
// costs ~8 % of the fG stage, while the tiny latency-bound ones (GRU gates, heads, fills) gain ~15 %.
template <bool kEarly = true, typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = getenv("AF_NO_PDL") == nullptr;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && kEarly) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Helper kernels are multi-wave and bandwidth-bound: they must NOT release their dependents early (a persistent
// conv CTA scheduled early would take an SM's registers / shared memory away from this kernel's remaining waves and
// then idle), so they only wait; the implicit trigger at grid completion releases the next kernel.
__device__ __forceinline__ void pdl_sync() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline unsigned grid_for(long long total, int threads = kThreads) {
  long long g = (total + threads - 1) / threads;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

// The reference evaluates floor(action * (H - P)) on fp32 tensors (ACT/models/utils.py:42): fp32 multiply, fp32
// floor, truncating cast. Clamped so a malformed action can never index out of bounds.
__device__ __forceinline__ int coord_from_action(float a, int H, int P) {
  const float span = static_cast<float>(H - P);
  int c = static_cast<int>(floorf(a * span));
  c = max(0, min(c, H - P));
  return c;
}

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == 1) return fmaxf(x, 0.f);
  if (act == 2) return fminf(fmaxf(x, 0.f), 6.f);
  return x;
}

// ------------------------------------------------------------------------------------------------ crop
__device__ __forceinline__ float4 load4_maybe_unaligned(const float* p) {
  if ((reinterpret_cast<uintptr_t>(p) & 15u) == 0) return __ldg(reinterpret_cast<const float4*>(p));
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}

template <int ROWS>
__global__ void __launch_bounds__(kThreads)
crop_nchw_f32_vec4_kernel(const float* __restrict__ img, const float* __restrict__ action,
                          const int32_t* __restrict__ yx, float* __restrict__ out, int32_t* __restrict__ yx_out,
                          int N, int C, int H, int W, int P) {
  const int p4 = P >> 2;
  const int prow = P / ROWS;   // rows handled per "slot": thread covers rows r, r+prow, ...
  const long long total = static_cast<long long>(N) * C * prow * p4;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int px4 = static_cast<int>(idx % p4);
  long long t = idx / p4;
  const int py = static_cast<int>(t % prow);
  t /= prow;
  const int c = static_cast<int>(t % C);
  const int n = static_cast<int>(t / C);
  int y0, x0;
  if (yx != nullptr) {
    y0 = max(0, min(yx[2 * n], H - P));
    x0 = max(0, min(yx[2 * n + 1], W - P));
  } else {
    y0 = coord_from_action(action[2 * n], H, P);
    x0 = min(coord_from_action(action[2 * n + 1], H, P), W - P);
  }
  if (yx_out != nullptr && c == 0 && py == 0 && px4 == 0) {
    yx_out[2 * n] = y0;
    yx_out[2 * n + 1] = x0;
  }
  const float* src = img + ((static_cast<long long>(n) * C + c) * H + y0) * W + x0 + px4 * 4;
  float* dst = out + (static_cast<long long>(n) * C + c) * P * P + px4 * 4;
  float4 v[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) v[r] = load4_maybe_unaligned(src + static_cast<long long>(py + r * prow) * W);
#pragma unroll
  for (int r = 0; r < ROWS; ++r) *reinterpret_cast<float4*>(dst + static_cast<long long>(py + r * prow) * P) = v[r];
}

__global__ void __launch_bounds__(kThreads)
crop_nchw_f32_scalar_kernel(const float* __restrict__ img, const float* __restrict__ action,
                            const int32_t* __restrict__ yx, float* __restrict__ out, int32_t* __restrict__ yx_out,
                            int N, int C, int H, int W, int P) {
  const long long total = static_cast<long long>(N) * C * P * P;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int px = static_cast<int>(idx % P);
  long long t = idx / P;
  const int py = static_cast<int>(t % P);
  t /= P;
  const int c = static_cast<int>(t % C);
  const int n = static_cast<int>(t / C);
  int y0, x0;
  if (yx != nullptr) {
    y0 = max(0, min(yx[2 * n], H - P));
    x0 = max(0, min(yx[2 * n + 1], W - P));
  } else {
    y0 = coord_from_action(action[2 * n], H, P);
    x0 = min(coord_from_action(action[2 * n + 1], H, P), W - P);
  }
  if (yx_out != nullptr && c == 0 && py == 0 && px == 0) {
    yx_out[2 * n] = y0;
    yx_out[2 * n + 1] = x0;
  }
  out[idx] = __ldg(img + ((static_cast<long long>(n) * C + c) * H + y0 + py) * W + x0 + px);
}

__global__ void action_to_yx_kernel(const float* __restrict__ action, int32_t* __restrict__ yx, int N, int H,
                                    int P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * N) yx[i] = coord_from_action(action[i], H, P);
}

// ------------------------------------------------------------------------------------------------ stem staging
// One block stages a strip of up to 64 output pixels of one output row: the (3 x KH x span) input window is read
// once, coalesced along x, converted to fp16 into shared memory (zero for conv padding outside the P x P patch), and
// the im2col rows are then assembled from shared memory and written as 16-byte chunks (coalesced along k).
constexpr int kStemStrip = 64;
constexpr int kStemMaxWindow = 3 * 7 * ((kStemStrip - 1) * 2 + 7);   // 3 ch x 7 rows x 133 cols
constexpr int kStemMaxK = 256;

__global__ void __launch_bounds__(kThreads)
stem_im2col_kernel(const float* __restrict__ frames, const int32_t* __restrict__ yx, int yx_div,
                   __half* __restrict__ out, int N, int H, int W, int P, int KH, int KW, int stride, int pad, int Ho,
                   int Wo, int Kpad) {
  pdl_sync();
  __shared__ __half s_win[kStemMaxWindow];
  __shared__ short s_off[kStemMaxK];
  const int n = blockIdx.z, oh = blockIdx.y, ow0 = blockIdx.x * kStemStrip;
  const int strip = min(kStemStrip, Wo - ow0);
  const int span = (strip - 1) * stride + KW;
  const int kreal = KH * KW * 3;
  int y0 = 0, x0 = 0;
  if (yx != nullptr) {   // one (y,x) per yx_div consecutive frames (STH: one crop per video division)
    const int e = n / yx_div;
    y0 = max(0, min(yx[2 * e], H - P));
    x0 = max(0, min(yx[2 * e + 1], W - P));
  }
  for (int k = threadIdx.x; k < Kpad; k += blockDim.x) {
    int off = -1;
    if (k < kreal) {
      const int tap = k / 3, c = k - tap * 3;
      const int kh = tap / KW, kw = tap - kh * KW;
      off = (c * KH + kh) * span + kw;
    }
    s_off[k] = static_cast<short>(off);
  }
  const float* base = frames + static_cast<long long>(n) * 3 * H * W;
  const int iy0 = oh * stride - pad, ix0 = ow0 * stride - pad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // window rows (c, r): one warp per row, lanes sweep x -> coalesced reads, no per-element div / mod
  for (int cr = warp; cr < 3 * KH; cr += nwarps) {
    const int c = cr / KH, r = cr - c * KH;
    const int iy = iy0 + r;
    const bool row_ok = (iy >= 0) && (iy < P);
    const float* rowp = base + (static_cast<long long>(c) * H + (y0 + iy)) * W + x0;
    __half* dst = s_win + cr * span;
    for (int xx = lane; xx < span; xx += 32) {
      const int ix = ix0 + xx;
      float v = 0.f;
      if (row_ok && ix >= 0 && ix < P) v = __ldg(rowp + ix);
      dst[xx] = __float2half_rn(v);
    }
  }
  __syncthreads();
  const int kgroups = Kpad >> 3;
  const long long row0 = (static_cast<long long>(n) * Ho + oh) * Wo + ow0;
  // one warp per output pixel, lane = 8-wide k group: a warp writes one contiguous im2col row
  for (int kg = lane; kg < kgroups; kg += 32) {
    short offs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) offs[j] = s_off[kg * 8 + j];
    for (int owl = warp; owl < strip; owl += nwarps) {
      __align__(16) __half vals[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) vals[j] = offs[j] >= 0 ? s_win[offs[j] + owl * stride] : __half(0.f);
      *reinterpret_cast<uint4*>(out + (row0 + owl) * Kpad + kg * 8) = *reinterpret_cast<const uint4*>(vals);
    }
  }
}

// ------------------------------------------------------------------------------------------------ crop + 2x2 space-to-depth
// A stride-2 KxK convolution over 3 channels equals a stride-1 ceil(K/2) x ceil(K/2) convolution over the 12 channels
// of the 2x2 space-to-depth image of the zero-padded input.  This kernel cuts the patch at (y0, x0) (get_patch,
// ACT/models/utils.py:37-51), applies the conv's zero padding, and writes out[n][Y][X][16] fp16 with channel
// (dy*2+dx)*3 + c = padded[c][2Y+dy][2X+dx] (channels 12-15 zero).  The tensor-core conv then reads it through a TMA
// view whose "pixel" is 4 consecutive X positions (64 channels, pixel stride 32 B: overlapping windows), so the
// horizontal taps ride in the channel dimension and the im2col matrix is never written.
// VT = 2 additionally folds the vertical neighbour into the pixel: out[n][Y][X][v*16 + (dy*2+dx)*3 + c] =
// padded[c][2(Y+v)+dy][2X+dx], 32 channels (64 B) per pixel -- a 3x3/2 stem then needs a window of only two X positions
// and becomes a single-k-block 1x1 conv over the view (half the operand traffic of the VT = 1 form).
template <int VT>
__global__ void __launch_bounds__(kThreads)
stem_s2d_kernel(const float* __restrict__ frames, const int32_t* __restrict__ yx, int yx_div,
                __half* __restrict__ out, int N, int H, int W, int P, int pad, int Hs, int Ws) {
  // block = 32 x 8 pixels of one image (grid: x tiles, y tiles, image): no integer divisions
  pdl_sync();
  const int X = blockIdx.x * 32 + (threadIdx.x & 31);
  const int Y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int n = blockIdx.z;
  if (X < Ws && Y < Hs) {
    const long long idx = (static_cast<long long>(n) * Hs + Y) * Ws + X;
    int y0 = 0, x0 = 0;
    if (yx != nullptr) {
      const int e = n / yx_div;
      y0 = max(0, min(yx[2 * e], H - P));
      x0 = max(0, min(yx[2 * e + 1], W - P));
    }
    const float* base = frames + static_cast<long long>(n) * 3 * H * W;
    uint4* dst = reinterpret_cast<uint4*>(out + idx * (16 * VT));
#pragma unroll
    for (int v = 0; v < VT; ++v) {
      float val[12];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int iy = 2 * (Y + v) + dy - pad;
        const bool row_ok = iy >= 0 && iy < P;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int ix = 2 * X + dx - pad;
          const bool ok = row_ok && ix >= 0 && ix < P;
#pragma unroll
          for (int c = 0; c < 3; ++c)
            val[(dy * 2 + dx) * 3 + c] =
                ok ? __ldg(base + (static_cast<long long>(c) * H + (y0 + iy)) * W + x0 + ix) : 0.f;
        }
      }
      __align__(16) __half2 h[8];
#pragma unroll
      for (int j = 0; j < 6; ++j) h[j] = __floats2half2_rn(val[2 * j], val[2 * j + 1]);
      h[6] = h[7] = __floats2half2_rn(0.f, 0.f);
      dst[2 * v] = *reinterpret_cast<const uint4*>(&h[0]);
      dst[2 * v + 1] = *reinterpret_cast<const uint4*>(&h[4]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ uint8 frame ingest
// Stack -> ToTorchFormatTensor(div=True) -> GroupNormalize of the reference's loaders (ACT/ops/transforms.py:303-336,
// 64-77) on the device: in (B, HW, C) uint8 (C = 3T, frame-major RGB, as np.concatenate(axis=2) leaves it) ->
// out (B, C, HW) fp32 = ((u / 255) - mean[c % 3]) / std[c % 3], every step rounded to fp32 like the torch ops
// (division, subtraction, division: no FMA contraction is possible).  One block transposes a 128-pixel strip
// through shared memory: 16-byte coalesced reads, 512-byte coalesced channel rows out.
constexpr int kIngestPix = 128;
constexpr int kIngestMaxC = 96;
__global__ void __launch_bounds__(kThreads)
u8hwc_to_f32chw_norm_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int HW, int C, float m0, float m1,
                            float m2, float s0, float s1, float s2) {
  __shared__ __align__(16) uint8_t tile[kIngestPix * kIngestMaxC + 16];
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * kIngestPix;
  const int npix = min(kIngestPix, HW - p0);
  const long long src = (static_cast<long long>(b) * HW + p0) * C;
  const int nbytes = npix * C;
  if ((src & 15) == 0) {
    for (int i = threadIdx.x * 16; i < nbytes; i += blockDim.x * 16) {
      if (i + 16 <= nbytes) {
        *reinterpret_cast<uint4*>(tile + i) = __ldg(reinterpret_cast<const uint4*>(in + src + i));
      } else {
        for (int j = i; j < nbytes; ++j) tile[j] = in[src + j];
      }
    }
  } else {
    for (int i = threadIdx.x; i < nbytes; i += blockDim.x) tile[i] = in[src + i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int c = warp; c < C; c += nwarps) {
    const int rgb = c % 3;
    const float m = rgb == 0 ? m0 : (rgb == 1 ? m1 : m2);
    const float sd = rgb == 0 ? s0 : (rgb == 1 ? s1 : s2);
    float* dst = out + (static_cast<long long>(b) * C + c) * HW + p0;
    for (int p = lane; p < npix; p += 32) {
      const float x = __fdiv_rn(static_cast<float>(tile[p * C + c]), 255.f);
      dst[p] = __fdiv_rn(__fsub_rn(x, m), sd);
    }
  }
}

// ------------------------------------------------------------------------------------------------ direct 3x3/2 stem
// MobileNet-V2 features[0] (ACT/models/mobilenet.py:105): 3 -> 32 channels, 3x3, stride 2, pad 1, BN, ReLU6, straight
// from the fp32 NCHW frame to NHWC fp16.  K = 27 is too thin for a 64-wide MMA k-block, so this layer runs on the
// FMA pipes: one output pixel x 32 channels per thread, weights broadcast from shared memory.
constexpr int kStemC = 32;

__global__ void __launch_bounds__(kThreads, 2)
stem_conv3x3s2_kernel(const float* __restrict__ frames, const float* __restrict__ w27, const float* __restrict__ scale,
                      const float* __restrict__ bias, __half* __restrict__ out, int N, int H, int W, int Ho, int Wo,
                      int act) {
  pdl_sync();
  // two horizontally adjacent output pixels per thread: every weight vector fetched from shared memory feeds 8 FMAs
  __shared__ __align__(16) float s_w[27 * kStemC];
  __shared__ float s_scale[kStemC], s_bias[kStemC];
  for (int i = threadIdx.x; i < 27 * kStemC; i += blockDim.x) s_w[i] = w27[i];
  if (threadIdx.x < kStemC) {
    s_scale[threadIdx.x] = scale[threadIdx.x];
    s_bias[threadIdx.x] = bias[threadIdx.x];
  }
  __syncthreads();
  const unsigned wpairs = (static_cast<unsigned>(Wo) + 1) >> 1;
  const unsigned total = static_cast<unsigned>(N) * Ho * wpairs;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ow = static_cast<int>(idx % wpairs) * 2;
  const int oh = static_cast<int>((idx / wpairs) % Ho);
  const int n = static_cast<int>(idx / (wpairs * Ho));
  float acc0[kStemC], acc1[kStemC];
#pragma unroll
  for (int j = 0; j < kStemC; ++j) acc0[j] = acc1[j] = 0.f;
  const float* base = frames + static_cast<long long>(n) * 3 * H * W;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int iy = oh * 2 + r - 1;
    if (iy < 0 || iy >= H) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* rowp = base + (static_cast<long long>(c) * H + iy) * W;
      float xin[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const int ix = ow * 2 + q - 1;
        xin[q] = (ix >= 0 && ix < W) ? __ldg(rowp + ix) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const float4* wp = reinterpret_cast<const float4*>(s_w + ((r * 3 + q) * 3 + c) * kStemC);
        const float x0 = xin[q], x1 = xin[q + 2];
#pragma unroll
        for (int j4 = 0; j4 < kStemC / 4; ++j4) {
          const float4 w = wp[j4];
          acc0[4 * j4 + 0] = fmaf(x0, w.x, acc0[4 * j4 + 0]);
          acc0[4 * j4 + 1] = fmaf(x0, w.y, acc0[4 * j4 + 1]);
          acc0[4 * j4 + 2] = fmaf(x0, w.z, acc0[4 * j4 + 2]);
          acc0[4 * j4 + 3] = fmaf(x0, w.w, acc0[4 * j4 + 3]);
          acc1[4 * j4 + 0] = fmaf(x1, w.x, acc1[4 * j4 + 0]);
          acc1[4 * j4 + 1] = fmaf(x1, w.y, acc1[4 * j4 + 1]);
          acc1[4 * j4 + 2] = fmaf(x1, w.z, acc1[4 * j4 + 2]);
          acc1[4 * j4 + 3] = fmaf(x1, w.w, acc1[4 * j4 + 3]);
        }
      }
    }
  }
  __half* op = out + ((static_cast<long long>(n) * Ho + oh) * Wo + ow) * kStemC;
#pragma unroll
  for (int px = 0; px < 2; ++px) {
    if (ow + px >= Wo) break;
    const float* acc = px == 0 ? acc0 : acc1;
#pragma unroll
    for (int g = 0; g < kStemC / 8; ++g) {
      uint4 ov;
      __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = g * 8 + 2 * j;
        oh2[j] = __floats2half2_rn(act_apply(fmaf(acc[c], s_scale[c], s_bias[c]), act),
                                   act_apply(fmaf(acc[c + 1], s_scale[c + 1], s_bias[c + 1]), act));
      }
      reinterpret_cast<uint4*>(op + px * kStemC)[g] = ov;
    }
  }
}

// ------------------------------------------------------------------------------------------------ depthwise 3x3
// Each thread produces WO horizontally adjacent output pixels x 8 channels: the 3 x ((WO-1)*S+3) input window is
// loaded once (16-byte loads) and the 9 per-channel weights once, instead of 9 loads per output pixel.
template <int S, int WO>
__global__ void __launch_bounds__(kThreads)
dwconv3x3_kernel(const __half* __restrict__ in, const float* __restrict__ w9c, const float* __restrict__ scale,
                 const float* __restrict__ bias, __half* __restrict__ out, int N, int H, int W, int C, int Ho,
                 int Wo, int act) {
  pdl_sync();
  constexpr int COLS = (WO - 1) * S + 3;
  const unsigned c8n = static_cast<unsigned>(C) >> 3;
  const unsigned wblocks = (static_cast<unsigned>(Wo) + WO - 1) / WO;
  const unsigned total = static_cast<unsigned>(N) * Ho * wblocks * c8n;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const unsigned c8 = idx % c8n;
  unsigned t = idx / c8n;
  const int ow0 = static_cast<int>(t % wblocks) * WO;
  t /= wblocks;
  const int oh = static_cast<int>(t % Ho);
  const int n = static_cast<int>(t / Ho);
  const int c = static_cast<int>(c8) * 8;
  float acc[WO][8];
#pragma unroll
  for (int o = 0; o < WO; ++o)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[o][j] = 0.f;
  const int ix0 = ow0 * S - 1;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int iy = oh * S + kh - 1;
    if (iy < 0 || iy >= H) continue;
    const __half* rowp = in + (static_cast<long long>(n) * H + iy) * W * C + c;
    float xin[COLS][8];
#pragma unroll
    for (int cc = 0; cc < COLS; ++cc) {
      const int ix = ix0 + cc;
      uint4 xv = make_uint4(0u, 0u, 0u, 0u);
      if (ix >= 0 && ix < W) xv = __ldg(reinterpret_cast<const uint4*>(rowp + static_cast<long long>(ix) * C));
      const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(xh[j]);
        xin[cc][2 * j] = f.x;
        xin[cc][2 * j + 1] = f.y;
      }
    }
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w9c + (kh * 3 + kw) * C + c));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w9c + (kh * 3 + kw) * C + c + 4));
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int o = 0; o < WO; ++o)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[o][j] = fmaf(xin[o * S + kw][j], wv[j], acc[o][j]);
    }
  }
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c));
  const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + c + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
  const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const float bi[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  __half* orow = out + ((static_cast<long long>(n) * Ho + oh) * Wo + ow0) * C + c;
#pragma unroll
  for (int o = 0; o < WO; ++o) {
    if (ow0 + o >= Wo) break;
    uint4 ov;
    __half2* oh2 = reinterpret_cast<__half2*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      oh2[j] = __floats2half2_rn(act_apply(fmaf(acc[o][2 * j], sc[2 * j], bi[2 * j]), act),
                                 act_apply(fmaf(acc[o][2 * j + 1], sc[2 * j + 1], bi[2 * j + 1]), act));
    *reinterpret_cast<uint4*>(orow + static_cast<long long>(o) * C) = ov;
  }
}

// ------------------------------------------------------------------------------------------------ pooling
__global__ void __launch_bounds__(kThreads)
maxpool3x3s2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int N, int H, int W, int C, int Ho,
                    int Wo) {
  // grid (x: output-column pairs x 8-channel groups, y: output row, z: image); a thread produces two adjacent output
  // pixels of one 8-channel group from a 3 x 5 input window (the middle column feeds both): 15 loads for 2 outputs,
  // 32-bit index math only
  pdl_sync();
  const int c8n = C >> 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ((Wo + 1) >> 1) * c8n) return;
  const int c8 = idx % c8n;
  const int ow0 = (idx / c8n) * 2;
  const int oh = blockIdx.y, n = blockIdx.z;
  const __half2 ninf = __float2half2_rn(-65504.f);
  __align__(16) __half2 m0[4] = {ninf, ninf, ninf, ninf};
  __align__(16) __half2 m1[4] = {ninf, ninf, ninf, ninf};
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int iy = oh * 2 + kh - 1;
    if (iy < 0 || iy >= H) continue;
    const __half* rowp = in + (static_cast<long long>(n) * H + iy) * W * C + c8 * 8;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int ix = ow0 * 2 - 1 + j;
      if (ix < 0 || ix >= W) continue;
      const uint4 xv = __ldg(reinterpret_cast<const uint4*>(rowp + static_cast<long long>(ix) * C));
      const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (j <= 2) m0[q] = __hmax2(m0[q], xh[q]);
        if (j >= 2) m1[q] = __hmax2(m1[q], xh[q]);
      }
    }
  }
  __half* op = out + ((static_cast<long long>(n) * Ho + oh) * Wo + ow0) * C + c8 * 8;
  *reinterpret_cast<uint4*>(op) = *reinterpret_cast<const uint4*>(m0);
  if (ow0 + 1 < Wo) *reinterpret_cast<uint4*>(op + C) = *reinterpret_cast<const uint4*>(m1);
}

__global__ void __launch_bounds__(kThreads)
avgpool_kernel(const __half* __restrict__ in, float* __restrict__ out_f32, long long out_f32_stride,
               __half* __restrict__ out_f16, long long out_f16_stride, int N, int HW, int C) {
  pdl_sync();
  const int c8n = C >> 3;
  const long long total = static_cast<long long>(N) * c8n;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = static_cast<int>(idx % c8n);
  const int n = static_cast<int>(idx / c8n);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const __half* p = in + static_cast<long long>(n) * HW * C + c8 * 8;
  for (int i = 0; i < HW; ++i) {
    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(p + static_cast<long long>(i) * C));
    const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(xh[j]);
      acc[2 * j] += f.x;
      acc[2 * j + 1] += f.y;
    }
  }
  const float inv = 1.f / static_cast<float>(HW);
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] *= inv;
  if (out_f32 != nullptr) {
    float* o = out_f32 + static_cast<long long>(n) * out_f32_stride + c8 * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = acc[j];
  }
  if (out_f16 != nullptr) {
    __half* o = out_f16 + static_cast<long long>(n) * out_f16_stride + c8 * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = __float2half_rn(acc[j]);
  }
}

// ------------------------------------------------------------------------------------------------ layout
__global__ void __launch_bounds__(kThreads)
nhwc_f16_to_nchw_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, int N, int HW, int C) {
  // 32x32 smem transpose over (p, c) for one image per blockIdx.z
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 256 threads -> 8 rows per pass
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    tile[r][tx] = (p < HW && c < C) ? __half2float(in[(static_cast<long long>(n) * HW + p) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    if (p < HW && c < C) out[(static_cast<long long>(n) * C + c) * HW + p] = tile[tx][r];
  }
}

__global__ void __launch_bounds__(kThreads)
nchw_f32_to_nhwc_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, int N, int C, int HW, int Cpad) {
  const long long total = static_cast<long long>(N) * HW * Cpad;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % Cpad);
  const long long np = idx / Cpad;
  const int p = static_cast<int>(np % HW);
  const int n = static_cast<int>(np / HW);
  out[idx] = (c < C) ? __float2half_rn(__ldg(in + (static_cast<long long>(n) * C + c) * HW + p)) : __half(0.f);
}

// ------------------------------------------------------------------------------------------------ GRU gates
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(kThreads)
gru_gates_kernel(const float* __restrict__ xg, long long xg_stride, const float* __restrict__ hg,
                 const float* __restrict__ h_prev, float* __restrict__ h_new, __half* __restrict__ h_new_f16,
                 __half* __restrict__ hseq_f16, long long hseq_stride, float* __restrict__ hseq_f32,
                 long long hseq_f32_stride, int B, int Hd, int split) {
  pdl_sync();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(B) * Hd) return;
  const int j = static_cast<int>(idx % Hd);
  const int b = static_cast<int>(idx / Hd);
  const float* x = xg + static_cast<long long>(b) * xg_stride;
  const float* h = hg + static_cast<long long>(b) * 3 * Hd;
  const float r = sigmoidf_(x[j] + h[j]);
  const float z = sigmoidf_(x[Hd + j] + h[Hd + j]);
  const float nn = tanhf(x[2 * Hd + j] + r * h[2 * Hd + j]);
  const float hp = h_prev[idx];
  const float hn = (1.f - z) * nn + z * hp;
  h_new[idx] = hn;
  const __half hi = __float2half_rn(hn);
  if (split) {
    // split-precision operand rows [hi | lo | hi] (3*Hd wide): the next GEMM multiplies them with [W_hi | W_hi | W_lo]
    const __half lo = __float2half_rn(hn - __half2float(hi));
    if (h_new_f16 != nullptr) {
      __half* o = h_new_f16 + static_cast<long long>(b) * 3 * Hd + j;
      o[0] = hi;
      o[Hd] = lo;
      o[2 * Hd] = hi;
    }
    if (hseq_f16 != nullptr) {
      __half* o = hseq_f16 + static_cast<long long>(b) * hseq_stride + j;
      o[0] = hi;
      o[Hd] = lo;
      o[2 * Hd] = hi;
    }
  } else {
    if (h_new_f16 != nullptr) h_new_f16[idx] = hi;
    if (hseq_f16 != nullptr) hseq_f16[static_cast<long long>(b) * hseq_stride + j] = hi;
  }
  if (hseq_f32 != nullptr) hseq_f32[static_cast<long long>(b) * hseq_f32_stride + j] = hn;
}

// ------------------------------------------------------------------------------------------------ GRU sequence
// Persistent recurrent kernel for small batches: all T steps of a GRU in ONE launch.  CTA c owns hidden units
// [c*JB, (c+1)*JB): its 3*JB rows of W_hh (fp16) stay in shared memory for the whole sequence.  Per step a warp takes
// a batch row, holds h_{t-1} (fp32) in registers, forms the 3*JB dot products with warp-shuffle reductions, lanes
// 0..JB-1 apply the gate math (torch.nn.GRU, gates r,z,n) and publish h_t; a grid-wide barrier (atomic counter,
// all CTAs co-resident: cooperative launch) separates the steps, so the T-step loop never leaves the device.
constexpr int kGruJB = 8;
constexpr int kGruMaxK = 4;   // hidden size <= 1024: h row cached as 4 x 8 floats per lane

__device__ __forceinline__ void gru_grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

// SPLIT: w_hh rows are [W_hi | W_hi | W_lo] (3*Hd wide, the split-precision packing of the classifier head); the lo
// parts are kept in shared memory as well and h_t is emitted as [hi | lo | hi] rows for the following GEMM.
template <bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
gru_sequence_kernel(const float* __restrict__ xg, const __half* __restrict__ w_hh, const float* __restrict__ b_hh,
                    const float* __restrict__ h0, float* __restrict__ hbuf, __half* __restrict__ hseq_f16,
                    long long hseq_stride, float* __restrict__ h_out, unsigned int* __restrict__ counter, int B, int T,
                    int Hd) {
  extern __shared__ __align__(16) uint8_t gru_smem[];
  constexpr int kParts = SPLIT ? 2 : 1;
  __half* s_w = reinterpret_cast<__half*>(gru_smem);                                   // [parts][3][JB][Hd]
  float* s_b = reinterpret_cast<float*>(gru_smem + sizeof(__half) * kParts * 3 * kGruJB * Hd);  // [3][JB]
  const long long w_stride = SPLIT ? 3ll * Hd : Hd;
  const int unit0 = blockIdx.x * kGruJB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int chunks = Hd >> 3;   // 16-byte chunks per row
  for (int i = threadIdx.x; i < 3 * kGruJB * chunks; i += blockDim.x) {
    const int row = i / chunks, ch = i - row * chunks;
    const int g = row / kGruJB, j = row - g * kGruJB;
    const __half* src = w_hh + (static_cast<long long>(g) * Hd + unit0 + j) * w_stride;
    reinterpret_cast<uint4*>(s_w)[i] = __ldg(reinterpret_cast<const uint4*>(src) + ch);
    if (SPLIT)
      reinterpret_cast<uint4*>(s_w + 3 * kGruJB * Hd)[i] = __ldg(reinterpret_cast<const uint4*>(src + 2 * Hd) + ch);
  }
  if (threadIdx.x < 3 * kGruJB) {
    const int g = threadIdx.x / kGruJB, j = threadIdx.x - g * kGruJB;
    s_b[threadIdx.x] = b_hh[g * Hd + unit0 + j];
  }
  // h_{-1}: every CTA initialises its own slice of buffer 0, then the grid synchronises
  for (int i = threadIdx.x; i < B * kGruJB; i += blockDim.x) {
    const int b = i / kGruJB, j = i - b * kGruJB;
    hbuf[static_cast<long long>(b) * Hd + unit0 + j] = h0 ? h0[static_cast<long long>(b) * Hd + unit0 + j] : 0.f;
  }
  unsigned int epoch = 1;
  gru_grid_barrier(counter, epoch * gridDim.x);
  const int kper = Hd >> 8;   // 8-element groups per lane (Hd / 256)
  for (int t = 0; t < T; ++t) {
    const float* hprev = hbuf + static_cast<long long>(t & 1) * B * Hd;
    float* hnext = hbuf + static_cast<long long>((t + 1) & 1) * B * Hd;
    for (int b = warp; b < B; b += nwarps) {
      const float* hb = hprev + static_cast<long long>(b) * Hd;
      // h_{t-1} of this batch row, 8 consecutive elements per (lane, i); L2-coherent loads (ld.global.cg): the row
      // was written by other CTAs during this launch and must not be served from a stale L1 line
      float hv[kGruMaxK][8];
#pragma unroll
      for (int i = 0; i < kGruMaxK; ++i) {
        if (i < kper) {
          const int k = (i * 32 + lane) * 8;
          const float4 a0 = __ldcg(reinterpret_cast<const float4*>(hb + k));
          const float4 a1 = __ldcg(reinterpret_cast<const float4*>(hb + k + 4));
          hv[i][0] = a0.x; hv[i][1] = a0.y; hv[i][2] = a0.z; hv[i][3] = a0.w;
          hv[i][4] = a1.x; hv[i][5] = a1.y; hv[i][6] = a1.z; hv[i][7] = a1.w;
        }
      }
      float my_r = 0.f, my_z = 0.f, my_n = 0.f;
      for (int j = 0; j < kGruJB; ++j) {
        float ar = 0.f, az = 0.f, an = 0.f;
#pragma unroll
        for (int i = 0; i < kGruMaxK; ++i) {
          if (i < kper) {
            const int k = (i * 32 + lane) * 8;
            const uint4 wr = *reinterpret_cast<const uint4*>(s_w + (0 * kGruJB + j) * Hd + k);
            const uint4 wz = *reinterpret_cast<const uint4*>(s_w + (1 * kGruJB + j) * Hd + k);
            const uint4 wn = *reinterpret_cast<const uint4*>(s_w + (2 * kGruJB + j) * Hd + k);
            const __half2* pr = reinterpret_cast<const __half2*>(&wr);
            const __half2* pz = reinterpret_cast<const __half2*>(&wz);
            const __half2* pn = reinterpret_cast<const __half2*>(&wn);
            uint4 lr, lz, ln;
            if (SPLIT) {
              const __half* s_lo = s_w + 3 * kGruJB * Hd;
              lr = *reinterpret_cast<const uint4*>(s_lo + (0 * kGruJB + j) * Hd + k);
              lz = *reinterpret_cast<const uint4*>(s_lo + (1 * kGruJB + j) * Hd + k);
              ln = *reinterpret_cast<const uint4*>(s_lo + (2 * kGruJB + j) * Hd + k);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float2 fr = __half22float2(pr[q]), fz = __half22float2(pz[q]), fn = __half22float2(pn[q]);
              if (SPLIT) {
                const float2 gr = __half22float2(reinterpret_cast<const __half2*>(&lr)[q]);
                const float2 gz = __half22float2(reinterpret_cast<const __half2*>(&lz)[q]);
                const float2 gn = __half22float2(reinterpret_cast<const __half2*>(&ln)[q]);
                fr.x += gr.x; fr.y += gr.y; fz.x += gz.x; fz.y += gz.y; fn.x += gn.x; fn.y += gn.y;
              }
              ar = fmaf(fr.x, hv[i][2 * q], ar);
              ar = fmaf(fr.y, hv[i][2 * q + 1], ar);
              az = fmaf(fz.x, hv[i][2 * q], az);
              az = fmaf(fz.y, hv[i][2 * q + 1], az);
              an = fmaf(fn.x, hv[i][2 * q], an);
              an = fmaf(fn.y, hv[i][2 * q + 1], an);
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          ar += __shfl_xor_sync(0xffffffffu, ar, o);
          az += __shfl_xor_sync(0xffffffffu, az, o);
          an += __shfl_xor_sync(0xffffffffu, an, o);
        }
        if (lane == j) {
          my_r = ar;
          my_z = az;
          my_n = an;
        }
      }
      if (lane < kGruJB) {
        const int unit = unit0 + lane;
        const float* x = xg + (static_cast<long long>(b) * T + t) * 3 * Hd;
        const float r = sigmoidf_(x[unit] + my_r + s_b[lane]);
        const float z = sigmoidf_(x[Hd + unit] + my_z + s_b[kGruJB + lane]);
        const float nn = tanhf(x[2 * Hd + unit] + r * (my_n + s_b[2 * kGruJB + lane]));
        const float hn = (1.f - z) * nn + z * __ldcg(hb + unit);
        hnext[static_cast<long long>(b) * Hd + unit] = hn;
        __half* hs = hseq_f16 + (static_cast<long long>(b) * T + t) * hseq_stride + unit;
        const __half hi = __float2half_rn(hn);
        hs[0] = hi;
        if (SPLIT) {
          hs[Hd] = __float2half_rn(hn - __half2float(hi));
          hs[2 * Hd] = hi;
        }
        if (t == T - 1 && h_out != nullptr) h_out[static_cast<long long>(b) * Hd + unit] = hn;
      }
    }
    ++epoch;
    gru_grid_barrier(counter, epoch * gridDim.x);
  }
}

// ------------------------------------------------------------------------------------------------ policy heads
__global__ void __launch_bounds__(kThreads)
policy_head_kernel(const float* __restrict__ logits, long long logit_stride, int A, int grid_n, int rows, int H,
                   int P, int32_t* __restrict__ action_idx, float* __restrict__ action_yx, int32_t* __restrict__ yx) {
  pdl_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* x = logits + static_cast<long long>(warp) * logit_stride;
  // softmax exactly as exp(x - max) / sum, then argmax over the probabilities (first maximum wins).
  float m = -INFINITY;
  for (int i = lane; i < A; i += 32) m = fmaxf(m, x[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int i = lane; i < A; i += 32) s += expf(x[i] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float best = -1.f;
  int best_i = 0x7fffffff;
  for (int i = lane; i < A; i += 32) {
    const float p = expf(x[i] - m) / s;
    if (p > best) {
      best = p;
      best_i = i;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ob > best || (ob == best && oi < best_i)) {
      best = ob;
      best_i = oi;
    }
  }
  if (lane == 0) {
    const int iy = best_i / grid_n, ix = best_i % grid_n;
    // table entries are python doubles k/(n-1) rounded to fp32 (ACT/models/gfv_net.py:272-307)
    const float ay = static_cast<float>(static_cast<double>(iy) / static_cast<double>(grid_n - 1));
    const float ax = static_cast<float>(static_cast<double>(ix) / static_cast<double>(grid_n - 1));
    if (action_idx != nullptr) action_idx[warp] = best_i;
    if (action_yx != nullptr) {
      action_yx[2 * warp] = ay;
      action_yx[2 * warp + 1] = ax;
    }
    if (yx != nullptr) {
      yx[2 * warp] = coord_from_action(ay, H, P);
      yx[2 * warp + 1] = coord_from_action(ax, H, P);
    }
  }
}

__global__ void policy_head_continuous_kernel(const float* __restrict__ logits, long long logit_stride, int rows,
                                              int H, int P, float* __restrict__ action_yx,
                                              int32_t* __restrict__ yx) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 2) return;
  const int r = i >> 1, k = i & 1;
  const float a = sigmoidf_(logits[static_cast<long long>(r) * logit_stride + k]);
  if (action_yx != nullptr) action_yx[i] = a;
  if (yx != nullptr) yx[i] = coord_from_action(a, H, P);
}

// ------------------------------------------------------------------------------------------------ TSM / consensus
__global__ void __launch_bounds__(kThreads)
tsm_shift_kernel(const __half* __restrict__ in, __half* __restrict__ out, int NT, int T, int HW, int C, int fold) {
  pdl_sync();
  const int c8n = C >> 3;
  const long long total = static_cast<long long>(NT) * HW * c8n;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = static_cast<int>(idx % c8n);
  const long long fp = idx / c8n;
  const int p = static_cast<int>(fp % HW);
  const int f = static_cast<int>(fp / HW);
  const int t = f % T;
  const int c = c8 * 8;
  __align__(16) __half vals[8];
  const long long frame = static_cast<long long>(HW) * C;
  const __half* base = in + static_cast<long long>(f) * frame + static_cast<long long>(p) * C;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int cc = c + j;
    __half v;
    if (cc < fold) v = (t + 1 < T) ? base[frame + cc] : __half(0.f);                 // from the next frame
    else if (cc < 2 * fold) v = (t > 0) ? *(base - frame + cc) : __half(0.f);        // from the previous frame
    else v = base[cc];
    vals[j] = v;
  }
  *reinterpret_cast<uint4*>(out + static_cast<long long>(f) * frame + static_cast<long long>(p) * C + c) =
      *reinterpret_cast<const uint4*>(vals);
}

// fp32 NCHW form of the shift, for the public TemporalShift.shift() on reference-layout tensors (pure copy)
__global__ void __launch_bounds__(kThreads)
tsm_shift_nchw_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int NT, int T, int C, int HW,
                          int fold) {
  const long long total = static_cast<long long>(NT) * C * HW;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long chw = static_cast<long long>(C) * HW;
  const int f = static_cast<int>(idx / chw);
  const int c = static_cast<int>((idx % chw) / HW);
  const int t = f % T;
  float v;
  if (c < fold) v = (t + 1 < T) ? __ldg(in + idx + chw) : 0.f;
  else if (c < 2 * fold) v = (t > 0) ? __ldg(in + idx - chw) : 0.f;
  else v = __ldg(in + idx);
  out[idx] = v;
}

__global__ void consensus_avg_kernel(const float* __restrict__ in, const float* __restrict__ add,
                                     float* __restrict__ out, int B, int T, int C) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C, b = i / C;
  float s = 0.f;
  for (int t = 0; t < T; ++t) s += in[(static_cast<long long>(b) * T + t) * C + c];
  s /= static_cast<float>(T);
  if (add != nullptr) s += add[i];
  out[i] = s;
}

// ------------------------------------------------------------------------------------------------ metrics (f-4)
// top-k hits: rank of the target's logit inside its row (ties: lower index first), one warp per row.
// ACT/ops/utils.py:35-49 (accuracy): hits[i] counts rows with rank < ks[i].
__global__ void __launch_bounds__(kThreads)
topk_hits_kernel(const float* __restrict__ logits, long long stride, const long long* __restrict__ target, int rows,
                 int C, int k0, int k1, float* __restrict__ hits) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* x = logits + static_cast<long long>(warp) * stride;
  const int t = static_cast<int>(target[warp]);
  if (t < 0 || t >= C) return;
  const float xt = x[t];
  int above = 0;
  for (int j = lane; j < C; j += 32) {
    const float v = x[j];
    above += (v > xt) || (v == xt && j < t);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
  if (lane == 0) {
    if (above < k0) atomicAdd(hits, 1.f);
    if (above < k1) atomicAdd(hits + 1, 1.f);
  }
}

// row softmax (fp32), one warp per row: the probabilities cal_map ranks (ACT/ops/utils.py:76)
__global__ void __launch_bounds__(kThreads)
softmax_rows_kernel(const float* __restrict__ logits, long long stride, float* __restrict__ probs, int rows, int C) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* x = logits + static_cast<long long>(warp) * stride;
  float m = -INFINITY;
  for (int j = lane; j < C; j += 32) m = fmaxf(m, x[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int j = lane; j < C; j += 32) s += expf(x[j] - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  for (int j = lane; j < C; j += 32) probs[static_cast<long long>(warp) * C + j] = expf(x[j] - m) / s;
}

// Average precision of one class per block (ACT/ops/utils.py:68-88): for every positive sample i, precision at its
// rank = (#positives ranked at or above i) / (rank of i); ranks come from counting, not sorting (ties: lower index
// first, i.e. a stable descending sort).  labels: (N, L) int64, -1 = no label; class k is positive for sample i if any
// of its L labels equals k.
__global__ void __launch_bounds__(kThreads)
class_ap_kernel(const float* __restrict__ probs, const long long* __restrict__ labels, int N, int C, int L,
                float* __restrict__ ap) {
  const int k = blockIdx.x;
  __shared__ float s_sum[kThreads];
  __shared__ int s_cnt[kThreads];
  float sum = 0.f;
  int npos = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    bool pos = false;
    for (int l = 0; l < L; ++l) pos |= (labels[static_cast<long long>(i) * L + l] == k);
    if (!pos) continue;
    ++npos;
    const float pi = probs[static_cast<long long>(i) * C + k];
    int rank = 1, tp = 1;
    for (int j = 0; j < N; ++j) {
      const float pj = probs[static_cast<long long>(j) * C + k];
      if ((pj > pi) || (pj == pi && j < i)) {
        ++rank;
        bool pj_pos = false;
        for (int l = 0; l < L; ++l) pj_pos |= (labels[static_cast<long long>(j) * L + l] == k);
        tp += pj_pos;
      }
    }
    sum += static_cast<float>(tp) / static_cast<float>(rank);
  }
  s_sum[threadIdx.x] = sum;
  s_cnt[threadIdx.x] = npos;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
      s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) ap[k] = s_sum[0] / fmaxf(static_cast<float>(s_cnt[0]), 1.f);
}

__global__ void fill_f32_kernel(float* p, float v, long long n) {
  pdl_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// ------------------------------------------------------------------------------------------------ PIL-exact resampling
// One pass of Pillow's 8-bit resampler (libImaging/Resample.c, ImagingResampleHorizontal_8bpc / Vertical_8bpc) behind
// torchvision.transforms.Resize, i.e. the reference's GroupScale (ACT/ops/transforms.py:78-93): per output index the
// host-computed window (first input index, tap count) and 22-bit fixed-point weights; acc = 2^21 + sum(pixel * k),
// >> 22, clipped to [0, 255].  Integer arithmetic throughout -> bit-identical to Pillow.  The tables only cover the
// rows / columns the centre crop keeps (GroupCenterCrop, :37-43), so the crop is free.
//   horizontal: in (N, H, W, C) -> tmp (N, rows, OW, C), rows = source rows [row0, row0 + rows)
//   vertical  : tmp (row r = source row row0 + r) -> out (N, OH, OW, C)
// Frames stay separate: af_frames_u8_to_f32 over N "clips" of C channels writes (N, C, HW), which IS the (B, T*C, HW)
// layout of Stack() + ToTorchFormatTensor (np.concatenate(img_group, axis=2) then permute, :303-336).
// Horizontal pass: one block per (frame, source row): the row is staged in shared memory with coalesced 16-byte loads,
// then every thread forms 4 consecutive output bytes (different pixels / channels) and stores them as one word.
constexpr int kResampleMaxRowBytes = 12288;   // W * C of the source frame (e.g. 4096 x 3)
__global__ void __launch_bounds__(kThreads)
pil_resample_h_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ bounds,
                         const int32_t* __restrict__ kk, int ksize, int H, int W, int C, int rows, int OW, int row0) {
  __shared__ __align__(16) uint8_t srow[kResampleMaxRowBytes + 32];
  const int n = blockIdx.y, y = blockIdx.x;
  const int in_bytes = W * C;
  const uint8_t* src = in + (static_cast<long long>(n) * H + (row0 + y)) * in_bytes;
  const int mis = static_cast<int>(reinterpret_cast<uintptr_t>(src) & 15u);   // stage from the aligned address below
  const uint8_t* asrc = src - mis;
  for (int i = threadIdx.x * 16; i < in_bytes + mis; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(srow + i) = __ldg(reinterpret_cast<const uint4*>(asrc + i));   // may over-read < 16 B inside the allocation's 256-B granule
  __syncthreads();
  const uint8_t* row = srow + mis;
  const int out_bytes = OW * C;
  uint8_t* dst = out + (static_cast<long long>(n) * rows + y) * out_bytes;
  for (int j0 = threadIdx.x * 4; j0 < out_bytes; j0 += blockDim.x * 4) {
    uint32_t packed = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = j0 + e;
      if (j < out_bytes) {
        const int x = j / C, c = j - x * C;
        const int first = __ldg(bounds + 2 * x), taps = __ldg(bounds + 2 * x + 1);
        const int32_t* k = kk + x * ksize;
        int acc = 1 << 21;
        for (int t = 0; t < taps; ++t) acc += static_cast<int>(row[(first + t) * C + c]) * __ldg(k + t);
        acc >>= 22;
        acc = acc < 0 ? 0 : (acc > 255 ? 255 : acc);
        packed |= static_cast<uint32_t>(acc) << (8 * e);
      }
    }
    if (j0 + 4 <= out_bytes && (reinterpret_cast<uintptr_t>(dst + j0) & 3u) == 0) {
      *reinterpret_cast<uint32_t*>(dst + j0) = packed;
    } else {
      for (int e = 0; e < 4 && j0 + e < out_bytes; ++e) dst[j0 + e] = static_cast<uint8_t>(packed >> (8 * e));
    }
  }
}

// Vertical pass: a thread forms 4 consecutive bytes of an output row from word loads of the tap rows.
__global__ void __launch_bounds__(kThreads)
pil_resample_v_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int32_t* __restrict__ bounds,
                         const int32_t* __restrict__ kk, int ksize, int N, int rows, int row_bytes, int OH, int row0) {
  const int words = (row_bytes + 3) >> 2;
  const long long total = static_cast<long long>(N) * OH * words;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wj = static_cast<int>(idx % words);
  const long long r = idx / words;
  const int y = static_cast<int>(r % OH);
  const int n = static_cast<int>(r / OH);
  const int first = __ldg(bounds + 2 * y), taps = __ldg(bounds + 2 * y + 1);
  const int32_t* k = kk + y * ksize;
  const uint8_t* src = in + (static_cast<long long>(n) * rows + (first - row0)) * row_bytes + wj * 4;
  uint8_t* dst = out + (static_cast<long long>(n) * OH + y) * row_bytes + wj * 4;
  const bool whole = wj * 4 + 4 <= row_bytes && (row_bytes & 3) == 0;   // aligned word access (buffers are 256-B aligned)
  int acc[4] = {1 << 21, 1 << 21, 1 << 21, 1 << 21};
  for (int t = 0; t < taps; ++t) {
    const int kv = __ldg(k + t);
    const uint8_t* p = src + static_cast<long long>(t) * row_bytes;
    if (whole) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
      acc[0] += static_cast<int>(v & 255u) * kv;
      acc[1] += static_cast<int>((v >> 8) & 255u) * kv;
      acc[2] += static_cast<int>((v >> 16) & 255u) * kv;
      acc[3] += static_cast<int>(v >> 24) * kv;
    } else {
      for (int e = 0; e < 4 && wj * 4 + e < row_bytes; ++e) acc[e] += static_cast<int>(p[e]) * kv;
    }
  }
  uint32_t packed = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int a = acc[e] >> 22;
    a = a < 0 ? 0 : (a > 255 ? 255 : a);
    packed |= static_cast<uint32_t>(a) << (8 * e);
  }
  if (whole) {
    *reinterpret_cast<uint32_t*>(dst) = packed;
  } else {
    for (int e = 0; e < 4 && wj * 4 + e < row_bytes; ++e) dst[e] = static_cast<uint8_t>(packed >> (8 * e));
  }
}

// fp32 rows -> split-precision fp16 operand rows [hi | lo | hi] (3*cols wide): x = hi + lo to ~22 bits; a GEMM against
// weights packed as [W_hi | W_hi | W_lo] then accumulates x_hi W_hi + x_lo W_hi + x_hi W_lo in fp32 on the tensor core
// (the W_lo x_lo term, 2^-22 relative, is dropped).  Used by the classifier head (ACT/models/gfv_net.py:427-435).
__global__ void split3_f16_kernel(const float* __restrict__ in, long long in_stride, __half* __restrict__ out, int rows,
                                  int cols) {
  pdl_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(rows) * cols) return;
  const int c = static_cast<int>(i % cols);
  const long long r = i / cols;
  const float v = in[r * in_stride + c];
  const __half hi = __float2half_rn(v);
  __half* o = out + r * 3 * cols + c;
  o[0] = hi;
  o[cols] = __float2half_rn(v - __half2float(hi));
  o[2 * cols] = hi;
}
__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, long long n) {
  pdl_sync();
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}

}  // namespace

// ==================================================================================================== launchers
cudaError_t launch_crop_nchw_f32(const float* img, const float* action, const int32_t* yx, float* out,
                                 int32_t* yx_out, int N, int C, int H, int W, int P, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  if ((P & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    if ((P % 16) == 0) {
      const long long total = static_cast<long long>(N) * C * (P / 4) * (P / 4);
      crop_nchw_f32_vec4_kernel<4><<<grid_for(total), kThreads, 0, s>>>(img, action, yx, out, yx_out, N, C, H, W, P);
    } else {
      const long long total = static_cast<long long>(N) * C * P * (P / 4);
      crop_nchw_f32_vec4_kernel<1><<<grid_for(total), kThreads, 0, s>>>(img, action, yx, out, yx_out, N, C, H, W, P);
    }
  } else {
    const long long total = static_cast<long long>(N) * C * P * P;
    crop_nchw_f32_scalar_kernel<<<grid_for(total), kThreads, 0, s>>>(img, action, yx, out, yx_out, N, C, H, W, P);
  }
  return cudaGetLastError();
}

cudaError_t launch_action_to_yx(const float* action, int32_t* yx, int N, int H, int P, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  action_to_yx_kernel<<<grid_for(2LL * N), kThreads, 0, s>>>(action, yx, N, H, P);
  return cudaGetLastError();
}

cudaError_t launch_stem_im2col(const float* frames, const int32_t* yx, int yx_div, __half* out, int N, int H, int W,
                               int P, int KH, int KW, int stride, int pad, int Ho, int Wo, int Kpad, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  if (KH > 7 || KW > 7 || stride > 2 || Kpad > kStemMaxK || N > 65535 || Ho > 65535) return cudaErrorInvalidValue;
  dim3 grid((Wo + kStemStrip - 1) / kStemStrip, Ho, N);
  return launch_pdl<false>(stem_im2col_kernel, dim3(grid), dim3(kThreads), 0, s, frames, yx, yx_div < 1 ? 1 : yx_div, out, N, H, W, P, KH, KW, stride,
                                               pad, Ho, Wo, Kpad);
  return cudaGetLastError();
}

cudaError_t launch_stem_s2d(const float* frames, const int32_t* yx, int yx_div, __half* out, int N, int H, int W, int P,
                            int pad, int Hs, int Ws, int vt, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  if (N > 65535) return cudaErrorInvalidValue;
  const dim3 grid((Ws + 31) / 32, (Hs + 7) / 8, N);
  if (vt == 2)
    return launch_pdl<false>(stem_s2d_kernel<2>, grid, dim3(kThreads), 0, s, frames, yx, yx_div < 1 ? 1 : yx_div, out,
                             N, H, W, P, pad, Hs, Ws);
  return launch_pdl<false>(stem_s2d_kernel<1>, grid, dim3(kThreads), 0, s, frames, yx, yx_div < 1 ? 1 : yx_div, out, N,
                           H, W, P, pad, Hs, Ws);
}

cudaError_t launch_stem_conv3x3s2(const float* frames, const float* w27, const float* scale, const float* bias,
                                  __half* out, int N, int H, int W, int act, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = static_cast<long long>(N) * Ho * ((Wo + 1) / 2);
  if (total >= (1LL << 32)) return cudaErrorInvalidValue;
  return launch_pdl<false>(stem_conv3x3s2_kernel, dim3(grid_for(total)), dim3(kThreads), 0, s, frames, w27, scale, bias, out, N, H, W, Ho, Wo, act);
  return cudaGetLastError();
}

cudaError_t launch_dwconv3x3(const __half* in, const float* w9c, const float* scale, const float* bias, __half* out,
                             int N, int H, int W, int C, int stride, int act, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
  constexpr int WO = 4;
  const long long total = static_cast<long long>(N) * Ho * ((Wo + WO - 1) / WO) * (C / 8);
  if (total >= (1LL << 32)) return cudaErrorInvalidValue;
  if (stride == 1)
    return launch_pdl<false>(dwconv3x3_kernel<1, WO>, dim3(grid_for(total)), dim3(kThreads), 0, s, in, w9c, scale, bias, out, N, H, W, C, Ho, Wo, act);
  else
    return launch_pdl<false>(dwconv3x3_kernel<2, WO>, dim3(grid_for(total)), dim3(kThreads), 0, s, in, w9c, scale, bias, out, N, H, W, C, Ho, Wo, act);
  return cudaGetLastError();
}

cudaError_t launch_maxpool3x3s2(const __half* in, __half* out, int N, int H, int W, int C, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  if (N > 65535 || Ho > 65535) return cudaErrorInvalidValue;
  const int per_row = ((Wo + 1) / 2) * (C / 8);
  const int threads = per_row >= kThreads ? kThreads : (per_row + 31) / 32 * 32;
  dim3 grid((per_row + threads - 1) / threads, Ho, N);
  return launch_pdl<false>(maxpool3x3s2_kernel, grid, dim3(threads), 0, s, in, out, N, H, W, C, Ho, Wo);
  return cudaGetLastError();
}

cudaError_t launch_avgpool(const __half* in, float* out_f32, long long out_f32_stride, __half* out_f16,
                           long long out_f16_stride, int N, int HW, int C, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  const long long total = static_cast<long long>(N) * (C / 8);
  return launch_pdl(avgpool_kernel, dim3(grid_for(total)), dim3(kThreads), 0, s, in, out_f32, out_f32_stride, out_f16, out_f16_stride, N, HW, C);
  return cudaGetLastError();
}

cudaError_t launch_nhwc_f16_to_nchw_f32(const __half* in, float* out, int N, int HW, int C, cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  dim3 grid((C + 31) / 32, (HW + 31) / 32, N);
  nhwc_f16_to_nchw_f32_kernel<<<grid, kThreads, 0, s>>>(in, out, N, HW, C);
  return cudaGetLastError();
}

cudaError_t launch_nchw_f32_to_nhwc_f16(const float* in, __half* out, int N, int C, int HW, int Cpad,
                                        cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  const long long total = static_cast<long long>(N) * HW * Cpad;
  nchw_f32_to_nhwc_f16_kernel<<<grid_for(total), kThreads, 0, s>>>(in, out, N, C, HW, Cpad);
  return cudaGetLastError();
}

cudaError_t launch_gru_gates(const float* xg, long long xg_stride, const float* hg, const float* h_prev,
                             float* h_new, __half* h_new_f16, __half* hseq_f16, long long hseq_stride,
                             float* hseq_f32, long long hseq_f32_stride, int B, int Hd, int split, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  return launch_pdl(gru_gates_kernel, dim3(grid_for(static_cast<long long>(B) * Hd)), dim3(kThreads), 0, s, xg, xg_stride,
                    hg, h_prev, h_new, h_new_f16, hseq_f16, hseq_stride, hseq_f32, hseq_f32_stride, B, Hd, split);
}

cudaError_t launch_split3_f16(const float* in, long long in_stride, __half* out, int rows, int cols, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  return launch_pdl(split3_f16_kernel, dim3(grid_for(static_cast<long long>(rows) * cols)), dim3(kThreads), 0, s, in,
                    in_stride, out, rows, cols);
}

cudaError_t launch_gru_sequence(const float* xg, const __half* w_hh, const float* b_hh, const float* h0, float* hbuf,
                                __half* hseq_f16, long long hseq_stride, float* h_out, unsigned int* counter, int B,
                                int T, int Hd, int sm_count, int split, cudaStream_t s) {
  if (B <= 0 || T <= 0) return cudaSuccess;
  if (Hd % 256 != 0 || Hd / kGruJB > sm_count || Hd > 256 * kGruMaxK) return cudaErrorInvalidValue;
  const size_t smem = sizeof(__half) * (split ? 2 : 1) * 3 * kGruJB * Hd + sizeof(float) * 3 * kGruJB;
  auto* kern = split ? gru_sequence_kernel<true> : gru_sequence_kernel<false>;
  if (smem > 48 * 1024) {   // per-device attribute; this launch is rare (one per GRU sequence), so set it every time
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), s);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(Hd / kGruJB);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // guarantees co-residency of all CTAs (grid barrier)
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, xg, w_hh, b_hh, h0, hbuf, hseq_f16, hseq_stride, h_out, counter, B, T, Hd);
}

cudaError_t launch_policy_head(const float* logits, long long logit_stride, int A, int grid_n, int rows, int H,
                               int P, int32_t* action_idx, float* action_yx, int32_t* yx, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  return launch_pdl(policy_head_kernel, dim3(grid_for(static_cast<long long>(rows) * 32)), dim3(kThreads), 0, s, logits,
                    logit_stride, A, grid_n, rows, H, P, action_idx, action_yx, yx);
}

cudaError_t launch_policy_head_continuous(const float* logits, long long logit_stride, int rows, int H, int P,
                                          float* action_yx, int32_t* yx, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  return launch_pdl(policy_head_continuous_kernel, dim3(grid_for(2LL * rows)), dim3(kThreads), 0, s, logits, logit_stride, rows, H, P,
                                                                        action_yx, yx);
  return cudaGetLastError();
}

cudaError_t launch_tsm_shift(const __half* in, __half* out, int NT, int T, int HW, int C, int fold,
                             cudaStream_t s) {
  if (NT <= 0) return cudaSuccess;
  const long long total = static_cast<long long>(NT) * HW * (C / 8);
  return launch_pdl<false>(tsm_shift_kernel, dim3(grid_for(total)), dim3(kThreads), 0, s, in, out, NT, T, HW, C, fold);
  return cudaGetLastError();
}

cudaError_t launch_tsm_shift_nchw_f32(const float* in, float* out, int NT, int T, int C, int HW, int fold,
                                      cudaStream_t s) {
  if (NT <= 0) return cudaSuccess;
  const long long total = static_cast<long long>(NT) * C * HW;
  tsm_shift_nchw_f32_kernel<<<grid_for(total), kThreads, 0, s>>>(in, out, NT, T, C, HW, fold);
  return cudaGetLastError();
}

cudaError_t launch_consensus_avg(const float* in, const float* add, float* out, int B, int T, int C,
                                 cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  return launch_pdl(consensus_avg_kernel, dim3(grid_for(static_cast<long long>(B) * C)), dim3(kThreads), 0, s, in, add, out,
                    B, T, C);
}

cudaError_t launch_topk_hits(const float* logits, long long stride, const long long* target, int rows, int C, int k0,
                             int k1, float* hits, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  topk_hits_kernel<<<grid_for(static_cast<long long>(rows) * 32), kThreads, 0, s>>>(logits, stride, target, rows, C, k0,
                                                                                   k1, hits);
  return cudaGetLastError();
}

cudaError_t launch_softmax_rows(const float* logits, long long stride, float* probs, int rows, int C, cudaStream_t s) {
  if (rows <= 0) return cudaSuccess;
  softmax_rows_kernel<<<grid_for(static_cast<long long>(rows) * 32), kThreads, 0, s>>>(logits, stride, probs, rows, C);
  return cudaGetLastError();
}

cudaError_t launch_class_ap(const float* probs, const long long* labels, int N, int C, int L, float* ap,
                            cudaStream_t s) {
  if (C <= 0) return cudaSuccess;
  class_ap_kernel<<<C, kThreads, 0, s>>>(probs, labels, N, C, L, ap);
  return cudaGetLastError();
}

cudaError_t launch_fill_f32(float* p, float v, long long n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  return launch_pdl(fill_f32_kernel, dim3(grid_for(n)), dim3(kThreads), 0, s, p, v, n);
  return cudaGetLastError();
}
cudaError_t launch_u8hwc_to_f32chw_norm(const uint8_t* in, float* out, int B, int HW, int C, const float* mean3,
                                        const float* std3, cudaStream_t s) {
  if (B <= 0 || HW <= 0) return cudaSuccess;
  if (C < 1 || C > kIngestMaxC || B > 65535) return cudaErrorInvalidValue;
  dim3 grid((HW + kIngestPix - 1) / kIngestPix, B);
  u8hwc_to_f32chw_norm_kernel<<<grid, kThreads, 0, s>>>(in, out, HW, C, mean3[0], mean3[1], mean3[2], std3[0], std3[1],
                                                        std3[2]);
  return cudaGetLastError();
}

cudaError_t launch_pil_resize_crop_u8(const uint8_t* in, uint8_t* tmp, uint8_t* out, int N, int H, int W, int C,
                                      const int32_t* hbounds, const int32_t* hkk, int hks, int OW,
                                      const int32_t* vbounds, const int32_t* vkk, int vks, int OH, int row0, int rows,
                                      cudaStream_t s) {
  if (N <= 0) return cudaSuccess;
  if (W * C > kResampleMaxRowBytes || N > 65535) return cudaErrorInvalidValue;
  pil_resample_h_u8_kernel<<<dim3(rows, N), kThreads, 0, s>>>(in, tmp, hbounds, hkk, hks, H, W, C, rows, OW, row0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const int row_bytes = OW * C;
  const long long total = static_cast<long long>(N) * OH * ((row_bytes + 3) >> 2);
  pil_resample_v_u8_kernel<<<grid_for(total), kThreads, 0, s>>>(tmp, out, vbounds, vkk, vks, N, rows, row_bytes, OH, row0);
  return cudaGetLastError();
}

cudaError_t launch_f32_to_f16(const float* in, __half* out, long long n, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  return launch_pdl(f32_to_f16_kernel, dim3(grid_for(n)), dim3(kThreads), 0, s, in, out, n);
  return cudaGetLastError();
}

}  // namespace af
