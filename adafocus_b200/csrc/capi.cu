// C ABI (include/adafocus_b200.h): context, plan recorder, tensor-map construction, kernel dispatch.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/adafocus_b200.h"
#include "conv_gemm.cuh"
#include "dwconv_tma.cuh"
#include "kernels.cuh"
#include "gru_tc.cuh"
#include "mbconv_fused.cuh"
#include "mbconv_rows.cuh"
#include "stem_gemm.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int fail_cuda(cudaError_t e, const char* where) {
  g_last_error = std::string(where) + ": " + cudaGetErrorString(e);
  return AF_ERR_CUDA;
}

using Launch = std::function<cudaError_t(cudaStream_t)>;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

struct af_plan {
  int device = 0;
  std::vector<Launch> launches;
  std::vector<char> is_mark;       // launches[i] is a timing-mark event record (skipped under stream capture)
  std::vector<cudaEvent_t> marks;
  int kernel_launches = 0;
  // forward I/O of a whole-path plan (af_plan_bind_forward): the static buffers the recorded launches read / write
  void* in_buf = nullptr;
  void* scan_buf = nullptr;
  size_t in_bytes = 0, scan_bytes = 0;
  const float* logits_buf = nullptr;
  int rows = 0, row_stride = 0, classes = 0, steps = 0;
  size_t workspace_bytes = 0;
  ~af_plan() {
    for (cudaEvent_t e : marks) cudaEventDestroy(e);
  }
};

struct af_ctx {
  int device = 0;
  int sm_count = 0;
  bool recording = false;
  af_plan* current = nullptr;
  EncodeTiledFn encode_tiled = nullptr;
};

namespace {

// Launches always go to the context's device, whatever the caller's current device is (restored on exit).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) cudaSetDevice(device);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int dispatch(af_ctx* ctx, void* stream, const char* name, Launch fn) {
  if (ctx == nullptr) return fail(AF_ERR_INVALID, std::string(name) + ": null ctx");
  if (ctx->recording) {
    ctx->current->launches.push_back(std::move(fn));
    ctx->current->is_mark.push_back(0);
    ctx->current->kernel_launches += 1;
    return AF_OK;
  }
  DeviceGuard guard(ctx->device);
  cudaError_t e = fn(static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return fail_cuda(e, name);
  return AF_OK;
}

bool encode_map(af_ctx* ctx, CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                const cuuint64_t* strides_bytes, const cuuint32_t* box, std::string* err,
                CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = ctx->encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank),
                                 const_cast<void*>(base), dims, strides_bytes, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf),
             "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] base %p",
             static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
             box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
    *err = buf;
    return false;
  }
  return true;
}

int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Pick the 128-pixel output box (TW x TH x TN, powers of two) that covers the output with the fewest tiles;
// ties go to the wider box (longer contiguous runs per TMA row).
void choose_tile(int N, int Ho, int Wo, int* TW, int* TH, int* TN) {
  long long best_cost = -1;
  for (int tw = 128; tw >= 1; tw >>= 1) {
    for (int th = 128 / tw; th >= 1; th >>= 1) {
      const int tn = 128 / (tw * th);
      const long long cost = static_cast<long long>(ceil_div(Wo, tw)) * ceil_div(Ho, th) * ceil_div(N, tn);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        *TW = tw;
        *TH = th;
        *TN = tn;
      }
    }
  }
}

// Same, with the image-axis extent of the box restricted to divisors of T (temporal-shift mode: a tile must not
// straddle two clips).
void choose_tile_tsm(int N, int Ho, int Wo, int T, int* TW, int* TH, int* TN) {
  long long best_cost = -1;
  for (int tw = 128; tw >= 1; tw >>= 1) {
    for (int th = 128 / tw; th >= 1; th >>= 1) {
      const int tn = 128 / (tw * th);
      if (T % tn != 0) continue;
      const long long cost = static_cast<long long>(ceil_div(Wo, tw)) * ceil_div(Ho, th) * ceil_div(N, tn);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        *TW = tw;
        *TH = th;
        *TN = tn;
      }
    }
  }
}

bool tsm_fold_ok(int n, int h, int w, int cin, long long in_stride, int fold, int t) {
  return t >= 1 && n >= t && n % t == 0 && h >= 1 && w >= 1 && cin % af::kConvBlockK == 0 && fold >= 16 &&
         fold % 16 == 0 && 2 * fold <= cin && in_stride >= cin && in_stride % 8 == 0;
}

}  // namespace

extern "C" {

int af_version(void) { return AF_VERSION; }

int af_conv_tsm_supported(int n, int h, int w, int cin, int in_stride, int fold, int t) {
  return tsm_fold_ok(n, h, w, cin, in_stride, fold, t) ? 1 : 0;
}
const char* af_last_error(void) { return g_last_error.c_str(); }

int af_ctx_create(af_ctx** out, int device) {
  if (out == nullptr) return fail(AF_ERR_INVALID, "af_ctx_create: null out");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return fail_cuda(e, "af_ctx_create: cudaGetDeviceCount (no CPU fallback exists)");
  if (device < 0 || device >= count) return fail(AF_ERR_INVALID, "af_ctx_create: bad device index");
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail_cuda(e, "af_ctx_create: cudaGetDeviceProperties");
  if (prop.major != 10) {
    return fail(AF_ERR_CUDA, "af_ctx_create: device is not sm_100 (Blackwell B200); this library has no other path");
  }
  af_ctx* ctx = new af_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    delete ctx;
    return fail(AF_ERR_CUDA, "af_ctx_create: cuTensorMapEncodeTiled entry point not available");
  }
  ctx->encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  *out = ctx;
  return AF_OK;
}

int af_ctx_destroy(af_ctx* ctx) {
  if (ctx == nullptr) return AF_OK;
  if (ctx->current != nullptr) delete ctx->current;
  delete ctx;
  return AF_OK;
}

int af_ctx_sm_count(const af_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

int af_plan_begin(af_ctx* ctx) {
  if (ctx == nullptr) return fail(AF_ERR_INVALID, "af_plan_begin: null ctx");
  if (ctx->recording) return fail(AF_ERR_STATE, "af_plan_begin: already recording");
  ctx->current = new af_plan();
  ctx->current->device = ctx->device;
  ctx->recording = true;
  return AF_OK;
}

int af_plan_end(af_ctx* ctx, af_plan** out) {
  if (ctx == nullptr || out == nullptr) return fail(AF_ERR_INVALID, "af_plan_end: null argument");
  if (!ctx->recording) return fail(AF_ERR_STATE, "af_plan_end: not recording");
  *out = ctx->current;
  ctx->current = nullptr;
  ctx->recording = false;
  return AF_OK;
}

int af_plan_run(af_plan* plan, void* stream) {
  if (plan == nullptr) return fail(AF_ERR_INVALID, "af_plan_run: null plan");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(plan->device);
  // Replays are CUDA-graph capturable (cudaStreamBeginCapture ... af_plan_run ... EndCapture): the kernel launches
  // (PDL edges included) become graph nodes; the timing marks are left out of a captured replay.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cap) != cudaSuccess) cap = cudaStreamCaptureStatusNone;
  for (size_t i = 0; i < plan->launches.size(); ++i) {
    if (cap != cudaStreamCaptureStatusNone && plan->is_mark[i]) continue;
    cudaError_t e = plan->launches[i](s);
    if (e != cudaSuccess) {
      char buf[64];
      snprintf(buf, sizeof(buf), "af_plan_run: launch %zu", i);
      return fail_cuda(e, buf);
    }
  }
  return AF_OK;
}

int af_plan_num_launches(const af_plan* plan) { return plan ? plan->kernel_launches : 0; }

int af_plan_mark(af_ctx* ctx, int* mark_index) {
  if (ctx == nullptr || !ctx->recording) return fail(AF_ERR_STATE, "af_plan_mark: only valid while recording");
  DeviceGuard guard(ctx->device);
  cudaEvent_t ev;
  cudaError_t e = cudaEventCreate(&ev);
  if (e != cudaSuccess) return fail_cuda(e, "af_plan_mark: cudaEventCreate");
  af_plan* plan = ctx->current;
  plan->marks.push_back(ev);
  if (mark_index != nullptr) *mark_index = static_cast<int>(plan->marks.size()) - 1;
  plan->launches.push_back([ev](cudaStream_t s) { return cudaEventRecord(ev, s); });
  plan->is_mark.push_back(1);
  return AF_OK;
}

int af_plan_mark_elapsed_ms(af_plan* plan, int mark_a, int mark_b, float* ms) {
  if (plan == nullptr || ms == nullptr || mark_a < 0 || mark_b < 0 ||
      mark_a >= static_cast<int>(plan->marks.size()) || mark_b >= static_cast<int>(plan->marks.size()))
    return fail(AF_ERR_INVALID, "af_plan_mark_elapsed_ms: bad mark index");
  cudaError_t e = cudaEventElapsedTime(ms, plan->marks[mark_a], plan->marks[mark_b]);
  if (e != cudaSuccess) return fail_cuda(e, "af_plan_mark_elapsed_ms (synchronise the stream first)");
  return AF_OK;
}

int af_plan_destroy(af_plan* plan) {
  delete plan;
  return AF_OK;
}

int af_plan_bind_forward(af_plan* plan, void* input_buf, size_t input_bytes, void* scan_buf, size_t scan_bytes,
                         const float* logits_buf, int rows, int row_stride, int classes, int steps,
                         size_t workspace_bytes) {
  if (plan == nullptr || input_buf == nullptr || logits_buf == nullptr || rows < 1 || classes < 1 ||
      row_stride < classes || steps < 1 || rows % steps != 0)
    return fail(AF_ERR_INVALID, "af_plan_bind_forward: bad argument");
  plan->in_buf = input_buf;
  plan->in_bytes = input_bytes;
  plan->scan_buf = scan_buf;
  plan->scan_bytes = scan_buf ? scan_bytes : 0;
  plan->logits_buf = logits_buf;
  plan->rows = rows;
  plan->row_stride = row_stride;
  plan->classes = classes;
  plan->steps = steps;
  plan->workspace_bytes = workspace_bytes;
  return AF_OK;
}

size_t af_workspace_bytes(const af_plan* plan) { return plan ? plan->workspace_bytes : 0; }

int af_gfv_forward(af_plan* plan, const float* input, const float* scan, float* logits, float* last_out, void* stream) {
  if (plan == nullptr || plan->in_buf == nullptr)
    return fail(AF_ERR_STATE, "af_gfv_forward: the plan has no forward binding (af_plan_bind_forward)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DeviceGuard guard(plan->device);
  cudaError_t e = cudaSuccess;
  if (input != nullptr && input != plan->in_buf)
    e = cudaMemcpyAsync(plan->in_buf, input, plan->in_bytes, cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return fail_cuda(e, "af_gfv_forward: input copy");
  if (plan->scan_buf != nullptr && scan != nullptr && scan != plan->scan_buf)
    e = cudaMemcpyAsync(plan->scan_buf, scan, plan->scan_bytes, cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return fail_cuda(e, "af_gfv_forward: scan copy");
  const int rc = af_plan_run(plan, stream);
  if (rc != AF_OK) return rc;
  const size_t w = static_cast<size_t>(plan->classes) * sizeof(float);
  const size_t pitch = static_cast<size_t>(plan->row_stride) * sizeof(float);
  if (logits != nullptr) {     // (B*T, C) contiguous, ACT/models/gfv_net.py:433
    e = cudaMemcpy2DAsync(logits, w, plan->logits_buf, pitch, w, plan->rows, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return fail_cuda(e, "af_gfv_forward: logits copy");
  }
  if (last_out != nullptr) {   // logits of the last step of every clip, :434
    const float* src = plan->logits_buf + static_cast<size_t>(plan->steps - 1) * plan->row_stride;
    e = cudaMemcpy2DAsync(last_out, w, src, pitch * plan->steps, w, plan->rows / plan->steps, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return fail_cuda(e, "af_gfv_forward: last_out copy");
  }
  return AF_OK;
}

// ------------------------------------------------------------------------------------------------ kernels
int af_crop_nchw_f32(af_ctx* ctx, const float* img, const float* action, const int32_t* yx, float* out,
                     int32_t* yx_out, int N, int C, int H, int W, int P, void* stream) {
  if (img == nullptr || out == nullptr) return fail(AF_ERR_INVALID, "af_crop_nchw_f32: null tensor");
  if ((action == nullptr) == (yx == nullptr))
    return fail(AF_ERR_INVALID, "af_crop_nchw_f32: exactly one of action / yx must be given");
  if (N < 0 || C <= 0 || P <= 0 || P > H || P > W)
    return fail(AF_ERR_INVALID, "af_crop_nchw_f32: need 0 < P <= min(H, W)");
  return dispatch(ctx, stream, "af_crop_nchw_f32", [=](cudaStream_t s) {
    return af::launch_crop_nchw_f32(img, action, yx, out, yx_out, N, C, H, W, P, s);
  });
}

int af_action_to_yx(af_ctx* ctx, const float* action, int32_t* yx, int N, int H, int P, void* stream) {
  if (action == nullptr || yx == nullptr || P > H) return fail(AF_ERR_INVALID, "af_action_to_yx: bad argument");
  return dispatch(ctx, stream, "af_action_to_yx",
                  [=](cudaStream_t s) { return af::launch_action_to_yx(action, yx, N, H, P, s); });
}

int af_stem_s2d(af_ctx* ctx, const float* frames, const int32_t* yx, int yx_div, void* out, int N, int H, int W, int P,
                int pad, int Hs, int Ws, int vt, void* stream) {
  if (frames == nullptr || out == nullptr) return fail(AF_ERR_INVALID, "af_stem_s2d: null tensor");
  if (P > H || P > W || P < 1 || pad < 0 || Hs < 1 || Ws < 1 || (vt != 1 && vt != 2))
    return fail(AF_ERR_INVALID, "af_stem_s2d: bad geometry");
  __half* o = static_cast<__half*>(out);
  return dispatch(ctx, stream, "af_stem_s2d", [=](cudaStream_t s) {
    return af::launch_stem_s2d(frames, yx, yx_div, o, N, H, W, P, pad, Hs, Ws, vt, s);
  });
}

int af_stem_im2col(af_ctx* ctx, const float* frames, const int32_t* yx, int yx_div, void* out, int N, int H, int W,
                   int P, int KH, int KW, int stride, int pad, int Kpad, void* stream) {
  if (frames == nullptr || out == nullptr) return fail(AF_ERR_INVALID, "af_stem_im2col: null tensor");
  if (Kpad % 8 != 0 || Kpad < KH * KW * 3 || P > H || P > W || stride < 1)
    return fail(AF_ERR_INVALID, "af_stem_im2col: bad geometry");
  const int Ho = (P + 2 * pad - KH) / stride + 1, Wo = (P + 2 * pad - KW) / stride + 1;
  __half* o = static_cast<__half*>(out);
  return dispatch(ctx, stream, "af_stem_im2col", [=](cudaStream_t s) {
    return af::launch_stem_im2col(frames, yx, yx_div, o, N, H, W, P, KH, KW, stride, pad, Ho, Wo, Kpad, s);
  });
}

int af_stem_conv_fused(af_ctx* ctx, const float* frames, const int32_t* yx, int yx_div, const void* w,
                       const float* scale, const float* bias, void* out, int N, int H, int W, int P, int cout, int KH,
                       int KW, int stride, int pad, int act, void* stream) {
  if (ctx == nullptr || frames == nullptr || w == nullptr || bias == nullptr || out == nullptr)
    return fail(AF_ERR_INVALID, "af_stem_conv_fused: null argument");
  if (P > H || P > W || N < 1) return fail(AF_ERR_INVALID, "af_stem_conv_fused: bad geometry");
  af::StemKernelParams p;
  memset(&p, 0, sizeof(p));
  p.frames = frames;
  p.yx = yx;
  p.yx_div = yx_div < 1 ? 1 : yx_div;
  p.N = N; p.H = H; p.W = W; p.P = P;
  p.KH = KH; p.KW = KW; p.stride = stride; p.pad = pad;
  p.Ho = (P + 2 * pad - KH) / stride + 1;
  p.Wo = (P + 2 * pad - KW) / stride + 1;
  if (p.Ho < 1 || p.Wo < 1) return fail(AF_ERR_INVALID, "af_stem_conv_fused: empty output");
  int tn = 1;
  choose_tile(1, p.Ho, p.Wo, &p.TW, &p.TH, &tn);
  if (tn != 1) return fail(AF_ERR_INVALID, "af_stem_conv_fused: output smaller than one 128-pixel tile");
  p.tiles_w = ceil_div(p.Wo, p.TW);
  p.tiles_h = ceil_div(p.Ho, p.TH);
  p.KB = ceil_div(KH * KW * 3, 64);
  p.BN = cout;
  p.scale = scale;
  p.bias = bias;
  p.act = act;
  if (!af::stem_gemm_supported(p))
    return fail(AF_ERR_INVALID, "af_stem_conv_fused: unsupported shape (cout multiple of 16 <= 64, K <= 256)");
  af::StemTensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  std::string err;
  {
    const cuuint64_t kpad = static_cast<cuuint64_t>(p.KB) * 64;
    const cuuint64_t dims[2] = {kpad, static_cast<cuuint64_t>(cout)};
    const cuuint64_t strides[1] = {kpad * 2};
    const cuuint32_t bbox[2] = {64, static_cast<cuuint32_t>(cout)};
    if (!encode_map(ctx, &maps.b, w, 2, dims, strides, bbox, &err)) return fail(AF_ERR_CUDA, err);
  }
  {
    const cuuint64_t pix_b = static_cast<cuuint64_t>(cout) * 2;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(cout), static_cast<cuuint64_t>(p.Wo),
                                static_cast<cuuint64_t>(p.Ho), static_cast<cuuint64_t>(N)};
    const cuuint64_t strides[3] = {pix_b, pix_b * p.Wo, pix_b * p.Wo * p.Ho};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(p.TW), static_cast<cuuint32_t>(p.TH), 1};
    if (!encode_map(ctx, &maps.out, out, 4, dims, strides, box, &err)) return fail(AF_ERR_CUDA, err);
  }
  const int sms = ctx->sm_count;
  return dispatch(ctx, stream, "af_stem_conv_fused",
                  [=](cudaStream_t s) { return af::launch_stem_gemm(maps, p, sms, s); });
}

int af_conv2d_nhwc_f16(af_ctx* ctx, const af_conv_desc* d, void* stream) {
  if (ctx == nullptr || d == nullptr) return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: null argument");
  if (d->in == nullptr || d->w == nullptr || d->out == nullptr || d->bias == nullptr)
    return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: null tensor");
  if (d->block_n < 16 || d->block_n > af::kConvMaxBlockN || d->block_n % 16 != 0)
    return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: block_n must be a multiple of 16 in [16,256]");
  if (d->stride != 1 && d->stride != 2) return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: stride must be 1 or 2");
  const bool windowed = d->in_row_stride != 0 || d->in_img_stride != 0;
  if (d->cin % 8 != 0 || d->in_stride % 8 != 0 || (!windowed && d->in_stride < d->cin))
    return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: cin and in_stride must be multiples of 8");
  if (windowed && (d->stride != 1 || d->in_row_stride % 8 != 0 || d->in_img_stride % 8 != 0 ||
                   d->in_row_stride < 8 || d->in_img_stride < 8))
    return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: explicit row/image strides need stride 1 and multiples of 8");
  if (d->cout < 1) return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: cout < 1");
  if (d->out_stride % 8 != 0 || (d->residual != nullptr && d->res_stride % 8 != 0))
    return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: out_stride / res_stride must be multiples of 8");
  if (d->n < 1 || d->h < 1 || d->w_ < 1 || d->kh < 1 || d->kw < 1)
    return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: bad geometry");
  if (d->stride == 2 && (d->h < 2 || d->w_ < 2))
    return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: stride-2 needs h, w >= 2");

  af::ConvKernelParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->n;
  p.Ho = (d->h + 2 * d->pad - d->kh) / d->stride + 1;
  p.Wo = (d->w_ + 2 * d->pad - d->kw) / d->stride + 1;
  if (p.Ho < 1 || p.Wo < 1) return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: empty output");
  p.Cout = d->cout;
  p.pair = af::conv_gemm_pair_ok(p.N, p.Ho, p.Wo, d->cout, d->block_n,
                                 d->kh * d->kw * ceil_div(d->cin, af::kConvBlockK) * af::kConvBlockK,
                                 d->residual != nullptr && d->scale == nullptr && !d->out_f32, ctx->sm_count) ? 1 : 0;
  const bool tsm = d->tsm_t > 0;
  if (tsm) {
    if (d->kh != 1 || d->kw != 1 || d->stride != 1 || d->pad != 0 || windowed ||
        !tsm_fold_ok(d->n, d->h, d->w_, d->cin, d->in_stride, d->tsm_fold, d->tsm_t))
      return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: temporal shift needs a dense 1x1 stride-1 conv, cin % 64 == 0, "
                                  "fold % 16 == 0 and n % tsm_t == 0 (see af_conv_tsm_supported)");
    choose_tile_tsm(p.N, p.Ho, p.Wo, d->tsm_t, &p.TW, &p.TH, &p.TN);
    p.tsm_T = d->tsm_t;
    p.tsm_f16 = d->tsm_fold / 16;
  } else {
    choose_tile(p.N, p.Ho, p.Wo, &p.TW, &p.TH, &p.TN);
  }
  {
    // vertical-halo mode (conv_gemm.cuh): stride-1 filters taller than one row whose weights stay resident; the tile
    // is one image's TW x TH pixels, chosen to minimise the rows loaded per tile (TH + KH - 1 per TH produced)
    static const bool no_vhalo = getenv("AF_NO_VHALO") != nullptr;
    const int nb = ceil_div(d->cout, d->block_n), cb = ceil_div(d->cin, af::kConvBlockK);
    if (!no_vhalo && d->stride == 1 && d->kh > 1 && af::conv_gemm_wres_ok(nb, d->block_n, d->kh, d->kw, cb, p.pair)) {
      long long best = -1;
      int btw = 0, bth = 0;
      for (int tw = 128; tw >= 8; tw >>= 1) {
        const int th = 128 / tw;
        if ((th + d->kh - 1) * tw * 128 > 48 * 1024) continue;
        const long long cost = 1LL * ceil_div(p.Wo, tw) * ceil_div(p.Ho, th) * (th + d->kh - 1) * tw;
        if (best < 0 || cost < best) {
          best = cost;
          btw = tw;
          bth = th;
        }
      }
      if (best > 0) {
        p.vhalo = 1;
        p.TW = btw;
        p.TH = bth;
        p.TN = 1;
      }
    }
  }
  if (d->pool) {
    // stem mode: MaxPool2d(3, 2, 1) in the epilogue -- tiles are two full output rows of one image
    const int nb = ceil_div(d->cout, d->block_n), cb = ceil_div(d->cin, af::kConvBlockK);
    if (d->stride != 1 || d->kh < 2 || p.Wo * 2 != af::kConvBlockM || (p.Ho & 1) || d->cout > 64 || d->cout % 8 != 0 ||
        d->out_f32 || d->residual != nullptr || d->act == AF_ACT_NONE || tsm ||
        !af::conv_gemm_wres_ok(nb, d->block_n, d->kh, d->kw, cb, 0))
      return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: pool needs a stride-1 multi-row filter with resident weights, "
                                  "a 64-pixel-wide even-height output, cout <= 64 and a ReLU-type activation");
    p.pool = 1;
    p.pair = 0;
    p.vhalo = 1;
    p.TW = p.Wo;
    p.TH = 2;
    p.TN = 1;
  }
  p.tiles_w = ceil_div(p.Wo, p.TW);
  p.tiles_h = ceil_div(p.Ho, p.TH);
  p.tiles_n = ceil_div(p.N, p.TN);
  p.BN = d->block_n;
  p.n_blocks = ceil_div(d->cout, d->block_n);
  p.KH = d->kh;
  p.KW = d->kw;
  p.stride = d->stride;
  p.pad = d->pad;
  p.cblks = ceil_div(d->cin, af::kConvBlockK);
  p.scale = d->scale;
  p.bias = d->bias;
  p.residual = static_cast<const __half*>(d->residual);
  p.res_stride = d->res_stride;
  p.out = d->out;
  p.out_stride = d->out_stride;
  p.out_f32 = d->out_f32;
  p.act = d->act;

  af::ConvTensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  std::string err;
  const __half* in = static_cast<const __half*>(d->in);
  const cuuint32_t box[4] = {static_cast<cuuint32_t>(af::kConvBlockK), static_cast<cuuint32_t>(p.TW),
                             static_cast<cuuint32_t>(p.TH), static_cast<cuuint32_t>(p.TN)};
  const cuuint64_t pix_b = static_cast<cuuint64_t>(d->in_stride) * 2;
  if (d->stride == 1) {
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->cin), static_cast<cuuint64_t>(d->w_),
                                static_cast<cuuint64_t>(d->h), static_cast<cuuint64_t>(d->n)};
    cuuint64_t strides[3] = {pix_b, pix_b * d->w_, pix_b * d->w_ * d->h};
    if (windowed) {
      strides[1] = static_cast<cuuint64_t>(d->in_row_stride) * 2;
      strides[2] = static_cast<cuuint64_t>(d->in_img_stride) * 2;
    }
    if (!encode_map(ctx, &maps.a[0], in, 4, dims, strides, box, &err)) return fail(AF_ERR_CUDA, err);
    if (tsm) {
      // {C, W, H, T, clips}: the frame axis is its own dimension so that t-1 / t+1 outside the clip are out of bounds
      const cuuint64_t dims5[5] = {static_cast<cuuint64_t>(d->cin), static_cast<cuuint64_t>(d->w_),
                                   static_cast<cuuint64_t>(d->h), static_cast<cuuint64_t>(d->tsm_t),
                                   static_cast<cuuint64_t>(d->n / d->tsm_t)};
      const cuuint64_t strides5[4] = {pix_b, pix_b * d->w_, pix_b * d->w_ * d->h, pix_b * d->w_ * d->h * d->tsm_t};
      const cuuint32_t box5[5] = {static_cast<cuuint32_t>(af::kConvBlockK), static_cast<cuuint32_t>(p.TW),
                                  static_cast<cuuint32_t>(p.TH), static_cast<cuuint32_t>(p.TN), 1};
      if (!encode_map(ctx, &maps.a5, in, 5, dims5, strides5, box5, &err)) return fail(AF_ERR_CUDA, err);
    }
    if (p.vhalo) {
      const cuuint32_t hbox[4] = {static_cast<cuuint32_t>(af::kConvBlockK), static_cast<cuuint32_t>(p.TW),
                                  static_cast<cuuint32_t>(p.TH + p.KH - 1), 1};
      if (!encode_map(ctx, &maps.ah, in, 4, dims, strides, hbox, &err)) return fail(AF_ERR_CUDA, err);
    }
  } else {
    for (int ph = 0; ph < 2; ++ph) {
      for (int pw = 0; pw < 2; ++pw) {
        const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->cin), static_cast<cuuint64_t>((d->w_ - pw + 1) / 2),
                                    static_cast<cuuint64_t>((d->h - ph + 1) / 2), static_cast<cuuint64_t>(d->n)};
        const cuuint64_t strides[3] = {pix_b * 2, pix_b * d->w_ * 2, pix_b * d->w_ * d->h};
        const __half* base = in + (static_cast<long long>(ph) * d->w_ + pw) * d->in_stride;
        if (!encode_map(ctx, &maps.a[ph * 2 + pw], base, 4, dims, strides, box, &err))
          return fail(AF_ERR_CUDA, err);
      }
    }
  }
  {
    const cuuint64_t kpad = static_cast<cuuint64_t>(d->kh) * d->kw * p.cblks * af::kConvBlockK;
    const cuuint64_t dims[2] = {kpad, static_cast<cuuint64_t>(p.n_blocks) * p.BN};
    const cuuint64_t strides[1] = {kpad * 2};
    // (CTA pairs: each CTA loads its own half of the tile's rows)
    const cuuint32_t bbox[2] = {static_cast<cuuint32_t>(af::kConvBlockK), static_cast<cuuint32_t>(p.pair ? p.BN / 2 : p.BN)};
    if (!encode_map(ctx, &maps.b, d->w, 2, dims, strides, bbox, &err)) return fail(AF_ERR_CUDA, err);
  }
  if (d->in2 != nullptr) {
    // second GEMM accumulated into the same tile: the projection shortcut downsample(x) of a bottleneck
    // (ACT/models/resnet.py:107-111), a 1x1 stride-s2 conv over the block input
    if (d->w2 == nullptr || d->cin2 < 8 || d->cin2 % 8 != 0 || d->in2_stride < d->cin2 || d->in2_stride % 8 != 0 ||
        (d->stride2 != 1 && d->stride2 != 2) || p.vhalo || tsm || p.pool || d->residual != nullptr || d->scale != nullptr)
      return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: in2/w2 need scale == NULL (folded), no residual, a non-halo conv, "
                                  "cin2 % 8 == 0 and stride2 in {1, 2}");
    const int h2 = d->stride2 == 1 ? p.Ho : d->h2, w2 = d->stride2 == 1 ? p.Wo : d->w2_;
    if ((h2 - 1) / d->stride2 + 1 != p.Ho || (w2 - 1) / d->stride2 + 1 != p.Wo)
      return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: in2 geometry does not produce the output grid");
    p.k2_blocks = ceil_div(d->cin2, af::kConvBlockK);
    const cuuint64_t pix2 = static_cast<cuuint64_t>(d->in2_stride) * 2;
    const cuuint64_t s2 = static_cast<cuuint64_t>(d->stride2);
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->cin2), static_cast<cuuint64_t>(p.Wo),
                                static_cast<cuuint64_t>(p.Ho), static_cast<cuuint64_t>(d->n)};
    const cuuint64_t strides[3] = {pix2 * s2, pix2 * w2 * s2, pix2 * w2 * h2};
    if (!encode_map(ctx, &maps.a2, d->in2, 4, dims, strides, box, &err)) return fail(AF_ERR_CUDA, err);
    const cuuint64_t kpad2 = static_cast<cuuint64_t>(p.k2_blocks) * af::kConvBlockK;
    const cuuint64_t bdims[2] = {kpad2, static_cast<cuuint64_t>(p.n_blocks) * p.BN};
    const cuuint64_t bstrides[1] = {kpad2 * 2};
    const cuuint32_t bbox2[2] = {static_cast<cuuint32_t>(af::kConvBlockK), static_cast<cuuint32_t>(p.pair ? p.BN / 2 : p.BN)};
    if (!encode_map(ctx, &maps.b2, d->w2, 2, bdims, bstrides, bbox2, &err)) return fail(AF_ERR_CUDA, err);
  }
  // fp16 outputs leave through the smem-staged TMA store when 64-channel slices never straddle an n-block
  p.tma_store = (!d->out_f32 && d->cout % 8 == 0 && (p.BN % 64 == 0 || p.n_blocks == 1)) ? 1 : 0;
  // A residual is added on the tensor core (R * I accumulated into TMEM), which needs the per-channel scale folded
  // into the weights (scale == NULL); with an explicit scale the slower direct-store epilogue adds it instead.
  if (d->residual != nullptr) {
    if (p.tma_store && d->scale == nullptr) p.res_mma = 1;
    else p.tma_store = 0;
  }
  if (p.pool) {
    // `out` is the POOLED tensor (N, Ho/2, Wo/2, cout); one pooled row of one image per TMA store
    if (!p.tma_store) return fail(AF_ERR_INVALID, "af_conv2d_nhwc_f16: pool needs the TMA-store epilogue");
    const cuuint64_t opix_b = static_cast<cuuint64_t>(d->out_stride) * 2;
    const cuuint64_t wp = static_cast<cuuint64_t>(p.Wo / 2), hp = static_cast<cuuint64_t>(p.Ho / 2);
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->cout), wp, hp, static_cast<cuuint64_t>(p.N)};
    const cuuint64_t strides[3] = {opix_b, opix_b * wp, opix_b * wp * hp};
    const cuuint32_t pbox[4] = {static_cast<cuuint32_t>(af::kConvBlockK), static_cast<cuuint32_t>(wp), 1, 1};
    if (!encode_map(ctx, &maps.pool, d->out, 4, dims, strides, pbox, &err)) return fail(AF_ERR_CUDA, err);
    maps.out = maps.pool;   // (only prefetched in this mode)
  } else if (p.tma_store) {
    const cuuint64_t opix_b = static_cast<cuuint64_t>(d->out_stride) * 2;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->cout), static_cast<cuuint64_t>(p.Wo),
                                static_cast<cuuint64_t>(p.Ho), static_cast<cuuint64_t>(p.N)};
    const cuuint64_t strides[3] = {opix_b, opix_b * p.Wo, opix_b * p.Wo * p.Ho};
    if (!encode_map(ctx, &maps.out, d->out, 4, dims, strides, box, &err)) return fail(AF_ERR_CUDA, err);
    if (p.res_mma) {
      const cuuint64_t rpix_b = static_cast<cuuint64_t>(d->res_stride) * 2;
      const cuuint64_t rstrides[3] = {rpix_b, rpix_b * p.Wo, rpix_b * p.Wo * p.Ho};
      if (!encode_map(ctx, &maps.res, d->residual, 4, dims, rstrides, box, &err)) return fail(AF_ERR_CUDA, err);
    }
  }
  const int sms = ctx->sm_count;
  return dispatch(ctx, stream, "af_conv2d_nhwc_f16",
                  [=](cudaStream_t s) { return af::launch_conv_gemm(maps, p, sms, s); });
}

namespace {

bool encode_mb_maps(af_ctx* ctx, const af_mbconv_desc* d, int BW, int BH, int TW, int TH, int Ho, int Wo, int nc,
                    int cout_pad, af::MbTensorMaps* maps, std::string* err) {
  {
    const cuuint64_t pix_b = static_cast<cuuint64_t>(d->cin) * 2;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->cin), static_cast<cuuint64_t>(d->w_),
                                static_cast<cuuint64_t>(d->h), static_cast<cuuint64_t>(d->n)};
    const cuuint64_t strides[3] = {pix_b, pix_b * d->w_, pix_b * d->w_ * d->h};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(BW), static_cast<cuuint32_t>(BH), 1};
    if (!encode_map(ctx, &maps->x, d->in, 4, dims, strides, box, err)) return false;
  }
  {
    const cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(nc) * 64};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, 64};
    if (!encode_map(ctx, &maps->w1, d->w1, 2, dims, strides, box, err)) return false;
  }
  {
    const cuuint64_t kpad = static_cast<cuuint64_t>(nc) * 64;
    const cuuint64_t dims[2] = {kpad, static_cast<cuuint64_t>(cout_pad)};
    const cuuint64_t strides[1] = {kpad * 2};
    const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(cout_pad)};
    if (!encode_map(ctx, &maps->w2, d->w2, 2, dims, strides, box, err)) return false;
  }
  {
    const cuuint64_t pix_b = static_cast<cuuint64_t>(d->cout) * 2;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->cout), static_cast<cuuint64_t>(Wo),
                                static_cast<cuuint64_t>(Ho), static_cast<cuuint64_t>(d->n)};
    const cuuint64_t strides[3] = {pix_b, pix_b * Wo, pix_b * Wo * Ho};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(TW), static_cast<cuuint32_t>(TH), 1};
    if (!encode_map(ctx, &maps->out, d->out, 4, dims, strides, box, err)) return false;
  }
  return true;
}

}  // namespace

int af_mbconv_fused_supported(int n, int h, int w, int cin, int cexp, int cout, int stride) {
  af::MbParams p;
  memset(&p, 0, sizeof(p));
  p.N = n; p.H = h; p.W = w; p.Cin = cin; p.Cexp = cexp; p.Cout = cout; p.S = stride;
  return af::mbconv_plan(&p) ? 1 : 0;
}

int af_mbconv_fused(af_ctx* ctx, const af_mbconv_desc* d, void* stream) {
  if (ctx == nullptr || d == nullptr) return fail(AF_ERR_INVALID, "af_mbconv_fused: null argument");
  if (d->in == nullptr || d->w1 == nullptr || d->bias1 == nullptr || d->dw_w == nullptr || d->bias2 == nullptr ||
      d->w2 == nullptr || d->bias3 == nullptr || d->out == nullptr)
    return fail(AF_ERR_INVALID, "af_mbconv_fused: null tensor");
  if (d->residual != nullptr && (d->stride != 1 || d->res_stride % 8 != 0 || d->res_stride < d->cout))
    return fail(AF_ERR_INVALID, "af_mbconv_fused: a residual needs stride 1 and res_stride % 8 == 0");
  af::MbTensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  std::string err;
  const int sms = ctx->sm_count;
  af::MbParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->n; p.H = d->h; p.W = d->w_; p.Cin = d->cin; p.Cexp = d->cexp; p.Cout = d->cout; p.S = d->stride;
  if (!af::mbconv_plan(&p)) return fail(AF_ERR_INVALID, "af_mbconv_fused: unsupported shape");
  p.bias_col = -1;
  if (d->bias1_in_w1) {
    if (d->cin + 2 > 64) return fail(AF_ERR_INVALID, "af_mbconv_fused: bias1_in_w1 needs cin + 2 <= 64");
    p.bias_col = d->cin;
    p.k1steps = (d->cin + 2 + 15) / 16;
  }
  p.bias1 = d->bias1; p.dw_w = d->dw_w; p.bias2 = d->bias2; p.bias3 = d->bias3;
  p.residual = static_cast<const __half*>(d->residual);
  p.res_stride = d->res_stride;
  if (!encode_mb_maps(ctx, d, p.BW, p.BH, p.TW, p.TH, p.Ho, p.Wo, p.nc, p.cout_pad, &maps, &err))
    return fail(AF_ERR_CUDA, err);
  return dispatch(ctx, stream, "af_mbconv_fused",
                  [=](cudaStream_t s) { return af::launch_mbconv_fused(maps, p, sms, s); });
}

static long long* g_mbrows_prof = nullptr;
/* bring-up hook: device buffer of 32 x 8 int64 for per-warp cycle counters of CTA 0 (instrumented builds only) */
int af_debug_mbconv_rows_prof(void* buf) {
  g_mbrows_prof = static_cast<long long*>(buf);
  return AF_OK;
}

int af_mbconv_rows_supported(int n, int h, int w, int cin, int cexp, int cout, int stride) {
  af::MrParams p;
  memset(&p, 0, sizeof(p));
  p.N = n; p.H = h; p.W = w; p.Cin = cin; p.Cexp = cexp; p.Cout = cout; p.S = stride;
  return af::mbrows_plan(&p) ? 1 : 0;
}

int af_mbconv_rows_layout(int cexp, int spr, int32_t* nchunks, int16_t* lane_ch, int16_t* lane_kpos) {
  if (nchunks == nullptr || lane_ch == nullptr || lane_kpos == nullptr)
    return fail(AF_ERR_INVALID, "af_mbconv_rows_layout: null argument");
  af::MrLayout lay;
  if (!af::mbrows_layout(cexp, spr, &lay)) return fail(AF_ERR_INVALID, "af_mbconv_rows_layout: unsupported (cexp, spr)");
  *nchunks = lay.nchunks;
  memcpy(lane_ch, lay.lane_ch, sizeof(lay.lane_ch));
  memcpy(lane_kpos, lay.lane_kpos, sizeof(lay.lane_kpos));
  return AF_OK;
}

int af_mbconv_rows(af_ctx* ctx, const af_mbconv_rows_desc* d, void* stream) {
  if (ctx == nullptr || d == nullptr) return fail(AF_ERR_INVALID, "af_mbconv_rows: null argument");
  if (d->in == nullptr || d->w1 == nullptr || d->dwp == nullptr || d->w2 == nullptr || d->bias3 == nullptr ||
      d->out == nullptr)
    return fail(AF_ERR_INVALID, "af_mbconv_rows: null tensor");
  if (d->residual != nullptr && (d->stride != 1 || d->res_stride % 8 != 0 || d->res_stride < d->cout))
    return fail(AF_ERR_INVALID, "af_mbconv_rows: a residual needs stride 1 and res_stride % 8 == 0");
  if (d->in_pix_stride % 8 != 0 || d->in_row_stride % 8 != 0 || d->in_img_stride % 8 != 0)
    return fail(AF_ERR_INVALID, "af_mbconv_rows: input view strides must be multiples of 8 elements");
  af::MrParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->n; p.H = d->h; p.W = d->w_; p.Cin = d->cin; p.Cexp = d->cexp; p.Cout = d->cout; p.S = d->stride;
  if (!af::mbrows_plan(&p)) return fail(AF_ERR_INVALID, "af_mbconv_rows: unsupported shape");
  p.dwp = d->dwp;
  p.bias3 = d->bias3;
  p.residual = static_cast<const __half*>(d->residual);
  p.res_stride = d->res_stride;
  p.prof = g_mbrows_prof;
  af::MrTensorMaps maps;
  memset(&maps, 0, sizeof(maps));
  std::string err;
  {
    const cuuint64_t pix_b = static_cast<cuuint64_t>(d->in_pix_stride > 0 ? d->in_pix_stride : d->cin) * 2;
    const cuuint64_t row_b = d->in_row_stride > 0 ? static_cast<cuuint64_t>(d->in_row_stride) * 2 : pix_b * d->w_;
    const cuuint64_t img_b = d->in_img_stride > 0 ? static_cast<cuuint64_t>(d->in_img_stride) * 2 : row_b * d->h;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->cin), static_cast<cuuint64_t>(d->w_),
                                static_cast<cuuint64_t>(d->h), static_cast<cuuint64_t>(d->n)};
    const cuuint64_t strides[3] = {pix_b, row_b, img_b};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(p.RP), static_cast<cuuint32_t>(p.G), 1};
    if (!encode_map(ctx, &maps.x, d->in, 4, dims, strides, box, &err)) return fail(AF_ERR_CUDA, err);
  }
  {
    const cuuint64_t dims[2] = {64, static_cast<cuuint64_t>(p.lay.nchunks) * 128};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {64, 128};
    if (!encode_map(ctx, &maps.w1, d->w1, 2, dims, strides, box, &err)) return fail(AF_ERR_CUDA, err);
  }
  {
    const cuuint64_t kpad = static_cast<cuuint64_t>(p.lay.nchunks) * 128;
    const cuuint64_t dims[2] = {kpad, static_cast<cuuint64_t>(p.cout_pad)};
    const cuuint64_t strides[1] = {kpad * 2};
    const cuuint32_t box[2] = {64, static_cast<cuuint32_t>(p.cout_pad)};
    if (!encode_map(ctx, &maps.w2, d->w2, 2, dims, strides, box, &err)) return fail(AF_ERR_CUDA, err);
  }
  {
    const cuuint64_t pix_b = static_cast<cuuint64_t>(d->cout) * 2;
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(d->cout), static_cast<cuuint64_t>(p.Wo),
                                static_cast<cuuint64_t>(p.Ho), static_cast<cuuint64_t>(d->n)};
    const cuuint64_t strides[3] = {pix_b, pix_b * p.Wo, pix_b * p.Wo * p.Ho};
    const cuuint32_t box[4] = {64, static_cast<cuuint32_t>(p.OWseg), static_cast<cuuint32_t>(p.OR * p.SPI), 1};
    if (!encode_map(ctx, &maps.out, d->out, 4, dims, strides, box, &err)) return fail(AF_ERR_CUDA, err);
    if (p.S == 1) {
      const cuuint32_t box_rest[4] = {64, static_cast<cuuint32_t>(p.OWseg), static_cast<cuuint32_t>(p.G - 1), 1};
      const cuuint32_t box_one[4] = {64, static_cast<cuuint32_t>(p.OWseg), 1, 1};
      if (!encode_map(ctx, &maps.out_rest, d->out, 4, dims, strides, box_rest, &err)) return fail(AF_ERR_CUDA, err);
      if (!encode_map(ctx, &maps.out_one, d->out, 4, dims, strides, box_one, &err)) return fail(AF_ERR_CUDA, err);
    }
  }
  const int sms = ctx->sm_count;
  return dispatch(ctx, stream, "af_mbconv_rows",
                  [=](cudaStream_t s) { return af::launch_mbconv_rows(maps, p, sms, s); });
}

int af_stem_conv3x3s2_c32(af_ctx* ctx, const float* frames, const float* w27, const float* scale, const float* bias,
                          void* out, int N, int H, int W, int act, void* stream) {
  if (frames == nullptr || w27 == nullptr || scale == nullptr || bias == nullptr || out == nullptr)
    return fail(AF_ERR_INVALID, "af_stem_conv3x3s2_c32: null tensor");
  if (H < 2 || W < 2) return fail(AF_ERR_INVALID, "af_stem_conv3x3s2_c32: bad geometry");
  __half* o = static_cast<__half*>(out);
  return dispatch(ctx, stream, "af_stem_conv3x3s2_c32", [=](cudaStream_t s) {
    return af::launch_stem_conv3x3s2(frames, w27, scale, bias, o, N, H, W, act, s);
  });
}

int af_dwconv3x3_nhwc_f16(af_ctx* ctx, const void* in, const float* w9c, const float* scale, const float* bias,
                          void* out, int N, int H, int W, int C, int stride, int act, void* stream) {
  if (in == nullptr || w9c == nullptr || scale == nullptr || bias == nullptr || out == nullptr)
    return fail(AF_ERR_INVALID, "af_dwconv3x3_nhwc_f16: null tensor");
  if (C % 8 != 0 || (stride != 1 && stride != 2)) return fail(AF_ERR_INVALID, "af_dwconv3x3_nhwc_f16: bad shape");
  const __half* i = static_cast<const __half*>(in);
  __half* o = static_cast<__half*>(out);
  static const bool direct_only = getenv("AF_DW_DIRECT") != nullptr;
  af::DwTmaParams tp;
  if (ctx != nullptr && !direct_only && af::dwconv_tma_plan(N, H, W, C, stride, &tp)) {
    // TMA-staged kernel: one 4-D box per tile in, one out (no swizzle: the threads read plain [y][x][c] tiles)
    tp.w9c = w9c; tp.scale = scale; tp.bias = bias; tp.act = act;
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    const int edge = stride == 1 ? 2 : 1;
    CUtensorMap in_map, out_map;
    std::string err;
    const cuuint64_t pix_b = static_cast<cuuint64_t>(C) * 2;
    {
      const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                                  static_cast<cuuint64_t>(N)};
      const cuuint64_t strides[3] = {pix_b, pix_b * W, pix_b * W * H};
      const cuuint32_t box[4] = {static_cast<cuuint32_t>(tp.CB), static_cast<cuuint32_t>(tp.TW * stride + edge),
                                 static_cast<cuuint32_t>(tp.TH * stride + edge), static_cast<cuuint32_t>(tp.NB)};
      if (!encode_map(ctx, &in_map, in, 4, dims, strides, box, &err, CU_TENSOR_MAP_SWIZZLE_NONE))
        return fail(AF_ERR_CUDA, err);
    }
    {
      const cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(Wo),
                                  static_cast<cuuint64_t>(Ho), static_cast<cuuint64_t>(N)};
      const cuuint64_t strides[3] = {pix_b, pix_b * Wo, pix_b * Wo * Ho};
      const cuuint32_t box[4] = {static_cast<cuuint32_t>(tp.CB), static_cast<cuuint32_t>(tp.TW),
                                 static_cast<cuuint32_t>(tp.TH), static_cast<cuuint32_t>(tp.NB)};
      if (!encode_map(ctx, &out_map, out, 4, dims, strides, box, &err, CU_TENSOR_MAP_SWIZZLE_NONE))
        return fail(AF_ERR_CUDA, err);
    }
    const int sms = ctx->sm_count;
    return dispatch(ctx, stream, "af_dwconv3x3_nhwc_f16", [=](cudaStream_t s) {
      return af::launch_dwconv3x3_tma(in_map, out_map, tp, stride, sms, s);
    });
  }
  return dispatch(ctx, stream, "af_dwconv3x3_nhwc_f16", [=](cudaStream_t s) {
    return af::launch_dwconv3x3(i, w9c, scale, bias, o, N, H, W, C, stride, act, s);
  });
}

int af_maxpool3x3s2_nhwc_f16(af_ctx* ctx, const void* in, void* out, int N, int H, int W, int C, void* stream) {
  if (in == nullptr || out == nullptr || C % 8 != 0) return fail(AF_ERR_INVALID, "af_maxpool3x3s2_nhwc_f16: bad argument");
  const __half* i = static_cast<const __half*>(in);
  __half* o = static_cast<__half*>(out);
  return dispatch(ctx, stream, "af_maxpool3x3s2_nhwc_f16",
                  [=](cudaStream_t s) { return af::launch_maxpool3x3s2(i, o, N, H, W, C, s); });
}

int af_avgpool_nhwc_f16(af_ctx* ctx, const void* in, float* out_f32, int64_t out_f32_stride, void* out_f16,
                        int64_t out_f16_stride, int N, int HW, int C, void* stream) {
  if (in == nullptr || (out_f32 == nullptr && out_f16 == nullptr) || C % 8 != 0 || HW < 1)
    return fail(AF_ERR_INVALID, "af_avgpool_nhwc_f16: bad argument");
  const __half* i = static_cast<const __half*>(in);
  __half* o16 = static_cast<__half*>(out_f16);
  return dispatch(ctx, stream, "af_avgpool_nhwc_f16", [=](cudaStream_t s) {
    return af::launch_avgpool(i, out_f32, out_f32_stride, o16, out_f16_stride, N, HW, C, s);
  });
}

int af_nhwc_f16_to_nchw_f32(af_ctx* ctx, const void* in, float* out, int N, int HW, int C, void* stream) {
  if (in == nullptr || out == nullptr) return fail(AF_ERR_INVALID, "af_nhwc_f16_to_nchw_f32: null tensor");
  if (N > 65535) return fail(AF_ERR_INVALID, "af_nhwc_f16_to_nchw_f32: N > 65535");
  const __half* i = static_cast<const __half*>(in);
  return dispatch(ctx, stream, "af_nhwc_f16_to_nchw_f32",
                  [=](cudaStream_t s) { return af::launch_nhwc_f16_to_nchw_f32(i, out, N, HW, C, s); });
}

int af_nchw_f32_to_nhwc_f16(af_ctx* ctx, const float* in, void* out, int N, int C, int HW, int Cpad, void* stream) {
  if (in == nullptr || out == nullptr || Cpad < C) return fail(AF_ERR_INVALID, "af_nchw_f32_to_nhwc_f16: bad argument");
  __half* o = static_cast<__half*>(out);
  return dispatch(ctx, stream, "af_nchw_f32_to_nhwc_f16",
                  [=](cudaStream_t s) { return af::launch_nchw_f32_to_nhwc_f16(in, o, N, C, HW, Cpad, s); });
}

int af_gru_gates(af_ctx* ctx, const float* xg, int64_t xg_stride, const float* hg, const float* h_prev, float* h_new,
                 void* h_new_f16, void* hseq_f16, int64_t hseq_stride, float* hseq_f32, int64_t hseq_f32_stride, int B,
                 int Hd, int split, void* stream) {
  if (xg == nullptr || hg == nullptr || h_prev == nullptr || h_new == nullptr)
    return fail(AF_ERR_INVALID, "af_gru_gates: null tensor");
  __half* a = static_cast<__half*>(h_new_f16);
  __half* b = static_cast<__half*>(hseq_f16);
  return dispatch(ctx, stream, "af_gru_gates", [=](cudaStream_t s) {
    return af::launch_gru_gates(xg, xg_stride, hg, h_prev, h_new, a, b, hseq_stride, hseq_f32, hseq_f32_stride, B, Hd,
                                split, s);
  });
}

int af_gru_sequence(af_ctx* ctx, const float* xg, const void* w_hh_f16, const float* b_hh, const float* h0, float* hbuf,
                    void* hseq_f16, int64_t hseq_stride, float* h_out, uint32_t* counter, int B, int T, int Hd,
                    int split, void* stream) {
  if (ctx == nullptr || xg == nullptr || w_hh_f16 == nullptr || b_hh == nullptr || hbuf == nullptr ||
      hseq_f16 == nullptr || counter == nullptr)
    return fail(AF_ERR_INVALID, "af_gru_sequence: null argument");
  if (Hd % 256 != 0 || Hd > 1024 || Hd / 8 > ctx->sm_count)
    return fail(AF_ERR_INVALID, "af_gru_sequence: hidden size must be a multiple of 256, <= 1024 and <= 8 * SM count");
  const __half* w = static_cast<const __half*>(w_hh_f16);
  __half* hs = static_cast<__half*>(hseq_f16);
  const int sms = ctx->sm_count;
  return dispatch(ctx, stream, "af_gru_sequence", [=](cudaStream_t s) {
    return af::launch_gru_sequence(xg, w, b_hh, h0, hbuf, hs, hseq_stride, h_out, counter, B, T, Hd, sms, split, s);
  });
}

int af_gru_sequence_tc_supported(const af_ctx* ctx, int B, int Hd, int split) {
  return (ctx != nullptr && af::gru_tc_supported(B, Hd, ctx->sm_count, split)) ? 1 : 0;
}

int af_gru_sequence_tc(af_ctx* ctx, const float* xg, const void* w_hh_f16, int64_t w_cols, const float* b_hh,
                       const float* h0, void* hbuf, void* hseq_f16, int64_t hseq_stride, float* h_out, uint32_t* counter,
                       int B, int T, int Hd, int split, void* stream) {
  if (ctx == nullptr || xg == nullptr || w_hh_f16 == nullptr || b_hh == nullptr || hbuf == nullptr ||
      hseq_f16 == nullptr || counter == nullptr)
    return fail(AF_ERR_INVALID, "af_gru_sequence_tc: null argument");
  if (!af::gru_tc_supported(B, Hd, ctx->sm_count, split) || T < 1 || w_cols < (split ? 3 : 1) * static_cast<int64_t>(Hd) ||
      hseq_stride < (split ? 3 : 1) * static_cast<int64_t>(Hd) || hseq_stride % 8 != 0)
    return fail(AF_ERR_INVALID, "af_gru_sequence_tc: needs 1 <= B <= 64, H % 64 == 0, H / 8 <= SM count "
                                "(af_gru_sequence_tc_supported)");
  af::GruTcMaps maps;
  memset(&maps, 0, sizeof(maps));
  std::string err;
  {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(w_cols), static_cast<cuuint64_t>(3 * Hd)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(w_cols) * 2};
    const cuuint32_t box[2] = {64, 8};
    if (!encode_map(ctx, &maps.w, w_hh_f16, 2, dims, strides, box, &err)) return fail(AF_ERR_CUDA, err);
  }
  const int kh = (split ? 2 : 1) * Hd;
  for (int i = 0; i < 2; ++i) {
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kh), static_cast<cuuint64_t>(B)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kh) * 2};
    const cuuint32_t box[2] = {64, 64};
    const __half* base = static_cast<const __half*>(hbuf) + static_cast<size_t>(i) * B * kh;
    if (!encode_map(ctx, &maps.h[i], base, 2, dims, strides, box, &err)) return fail(AF_ERR_CUDA, err);
  }
  af::GruTcParams p;
  memset(&p, 0, sizeof(p));
  p.xg = xg;
  p.b_hh = b_hh;
  p.h0 = h0;
  p.hbuf = static_cast<__half*>(hbuf);
  p.hseq = static_cast<__half*>(hseq_f16);
  p.hseq_stride = hseq_stride;
  p.h_out = h_out;
  p.counter = counter;
  p.B = B;
  p.T = T;
  p.H = Hd;
  return dispatch(ctx, stream, "af_gru_sequence_tc", [=](cudaStream_t s) { return af::launch_gru_tc(maps, p, split, s); });
}

int af_policy_head(af_ctx* ctx, const float* logits, int64_t logit_stride, int A, int grid_n, int rows, int H, int P,
                   int32_t* action_idx, float* action_yx, int32_t* yx, void* stream) {
  if (logits == nullptr || A < 1 || grid_n < 2 || grid_n * grid_n != A || P > H)
    return fail(AF_ERR_INVALID, "af_policy_head: need A == grid_n^2, grid_n >= 2, P <= H");
  return dispatch(ctx, stream, "af_policy_head", [=](cudaStream_t s) {
    return af::launch_policy_head(logits, logit_stride, A, grid_n, rows, H, P, action_idx, action_yx, yx, s);
  });
}

int af_policy_head_continuous(af_ctx* ctx, const float* logits, int64_t logit_stride, int rows, int H, int P,
                              float* action_yx, int32_t* yx, void* stream) {
  if (logits == nullptr || P > H) return fail(AF_ERR_INVALID, "af_policy_head_continuous: bad argument");
  return dispatch(ctx, stream, "af_policy_head_continuous", [=](cudaStream_t s) {
    return af::launch_policy_head_continuous(logits, logit_stride, rows, H, P, action_yx, yx, s);
  });
}

int af_tsm_shift_nhwc_f16(af_ctx* ctx, const void* in, void* out, int NT, int T, int HW, int C, int fold,
                          void* stream) {
  if (in == nullptr || out == nullptr || in == out || C % 8 != 0 || T < 1 || NT % T != 0 || 2 * fold > C)
    return fail(AF_ERR_INVALID, "af_tsm_shift_nhwc_f16: bad argument (NT must be a multiple of T, out != in)");
  const __half* i = static_cast<const __half*>(in);
  __half* o = static_cast<__half*>(out);
  return dispatch(ctx, stream, "af_tsm_shift_nhwc_f16",
                  [=](cudaStream_t s) { return af::launch_tsm_shift(i, o, NT, T, HW, C, fold, s); });
}

int af_tsm_shift_nchw_f32(af_ctx* ctx, const float* in, float* out, int NT, int T, int C, int HW, int fold,
                          void* stream) {
  if (in == nullptr || out == nullptr || in == out || T < 1 || NT % T != 0 || 2 * fold > C)
    return fail(AF_ERR_INVALID, "af_tsm_shift_nchw_f32: bad argument (NT must be a multiple of T, out != in)");
  return dispatch(ctx, stream, "af_tsm_shift_nchw_f32",
                  [=](cudaStream_t s) { return af::launch_tsm_shift_nchw_f32(in, out, NT, T, C, HW, fold, s); });
}

int af_consensus_avg(af_ctx* ctx, const float* in, const float* add, float* out, int B, int T, int C, void* stream) {
  if (in == nullptr || out == nullptr || T < 1) return fail(AF_ERR_INVALID, "af_consensus_avg: bad argument");
  return dispatch(ctx, stream, "af_consensus_avg",
                  [=](cudaStream_t s) { return af::launch_consensus_avg(in, add, out, B, T, C, s); });
}

int af_topk_hits(af_ctx* ctx, const float* logits, int64_t stride, const int64_t* target, int rows, int C, int k0,
                 int k1, float* hits, void* stream) {
  if (logits == nullptr || target == nullptr || hits == nullptr || C < 1)
    return fail(AF_ERR_INVALID, "af_topk_hits: bad argument");
  const long long* t = reinterpret_cast<const long long*>(target);
  return dispatch(ctx, stream, "af_topk_hits",
                  [=](cudaStream_t s) { return af::launch_topk_hits(logits, stride, t, rows, C, k0, k1, hits, s); });
}

int af_softmax_rows(af_ctx* ctx, const float* logits, int64_t stride, float* probs, int rows, int C, void* stream) {
  if (logits == nullptr || probs == nullptr || C < 1) return fail(AF_ERR_INVALID, "af_softmax_rows: bad argument");
  return dispatch(ctx, stream, "af_softmax_rows",
                  [=](cudaStream_t s) { return af::launch_softmax_rows(logits, stride, probs, rows, C, s); });
}

int af_class_ap(af_ctx* ctx, const float* probs, const int64_t* labels, int N, int C, int L, float* ap, void* stream) {
  if (probs == nullptr || labels == nullptr || ap == nullptr || N < 0 || C < 1 || L < 1)
    return fail(AF_ERR_INVALID, "af_class_ap: bad argument");
  const long long* l = reinterpret_cast<const long long*>(labels);
  return dispatch(ctx, stream, "af_class_ap",
                  [=](cudaStream_t s) { return af::launch_class_ap(probs, l, N, C, L, ap, s); });
}

int af_fill_f32(af_ctx* ctx, float* p, float v, int64_t n, void* stream) {
  if (p == nullptr) return fail(AF_ERR_INVALID, "af_fill_f32: null tensor");
  return dispatch(ctx, stream, "af_fill_f32", [=](cudaStream_t s) { return af::launch_fill_f32(p, v, n, s); });
}

int af_frames_u8_to_f32(af_ctx* ctx, const uint8_t* in, float* out, int B, int HW, int C, const float* mean3,
                        const float* std3, void* stream) {
  if (in == nullptr || out == nullptr || mean3 == nullptr || std3 == nullptr)
    return fail(AF_ERR_INVALID, "af_frames_u8_to_f32: null argument");
  if (B < 0 || HW < 1 || C < 1 || C > 96 || C % 3 != 0)
    return fail(AF_ERR_INVALID, "af_frames_u8_to_f32: C must be a multiple of 3 in [3, 96]");
  if (std3[0] == 0.f || std3[1] == 0.f || std3[2] == 0.f) return fail(AF_ERR_INVALID, "af_frames_u8_to_f32: zero std");
  const float m[3] = {mean3[0], mean3[1], mean3[2]}, sd[3] = {std3[0], std3[1], std3[2]};
  return dispatch(ctx, stream, "af_frames_u8_to_f32", [=](cudaStream_t s) {
    return af::launch_u8hwc_to_f32chw_norm(in, out, B, HW, C, m, sd, s);
  });
}

int af_split3_f16(af_ctx* ctx, const float* in, int64_t in_stride, void* out, int rows, int cols, void* stream) {
  if (in == nullptr || out == nullptr || in_stride < cols) return fail(AF_ERR_INVALID, "af_split3_f16: bad argument");
  __half* o = static_cast<__half*>(out);
  return dispatch(ctx, stream, "af_split3_f16",
                  [=](cudaStream_t s) { return af::launch_split3_f16(in, in_stride, o, rows, cols, s); });
}

int af_resize_crop_u8(af_ctx* ctx, const uint8_t* in, uint8_t* tmp, uint8_t* out, int N, int H, int W, int C,
                      const int32_t* hbounds, const int32_t* hkk, int hks, int OW, const int32_t* vbounds,
                      const int32_t* vkk, int vks, int OH, int row0, int rows, void* stream) {
  if (in == nullptr || tmp == nullptr || out == nullptr || hbounds == nullptr || hkk == nullptr || vbounds == nullptr ||
      vkk == nullptr)
    return fail(AF_ERR_INVALID, "af_resize_crop_u8: null argument");
  if (N < 0 || H < 1 || W < 1 || C < 1 || OW < 1 || OH < 1 || hks < 1 || vks < 1 || row0 < 0 || rows < 1 ||
      row0 + rows > H || W * C > 12288 || N > 65535)
    return fail(AF_ERR_INVALID, "af_resize_crop_u8: bad geometry (need row0 + rows <= H, W * C <= 12288, N <= 65535)");
  return dispatch(ctx, stream, "af_resize_crop_u8", [=](cudaStream_t s) {
    return af::launch_pil_resize_crop_u8(in, tmp, out, N, H, W, C, hbounds, hkk, hks, OW, vbounds, vkk, vks, OH, row0,
                                         rows, s);
  });
}

int af_f32_to_f16(af_ctx* ctx, const float* in, void* out, int64_t n, void* stream) {
  if (in == nullptr || out == nullptr) return fail(AF_ERR_INVALID, "af_f32_to_f16: null tensor");
  __half* o = static_cast<__half*>(out);
  return dispatch(ctx, stream, "af_f32_to_f16", [=](cudaStream_t s) { return af::launch_f32_to_f16(in, o, n, s); });
}

}  // extern "C"
