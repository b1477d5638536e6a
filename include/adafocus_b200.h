/*
 * adafocus_b200 -- C ABI of the B200 (sm_100a) AdaFocus offline-inference hot path.
 *
 * The reference (blackfeather-wang/AdaFocus @ 8c0f8d2) is pure Python/PyTorch and has no FFI; its boundary for
 * this path is the Python surface of models/ (GFV.forward / glance / get_patch ...).  This header is the native
 * boundary that the Python mirror in adafocus_b200/models/ binds with ctypes.  Each entry
 * point names the reference call site it replaces.  Aliases: ACT/ = "Experiments on ActivityNet, FCVID and
 * Mini-Kinetics/", STH/ = "Experiments on Something-Something V1&V2/".
 *
 * Conventions
 *   - every pointer is caller-owned DEVICE memory (torch tensors' data_ptr()); the library never frees or keeps
 *     them beyond the call, except that a recorded plan replays launches on the same pointers;
 *   - every call is asynchronous on the given stream (a cudaStream_t passed as void*), never synchronises, and is
 *     CUDA-graph capturable;
 *   - return 0 on success, a negative af_status otherwise; af_last_error() gives a thread-local message;
 *   - activations between layers are NHWC fp16 ("half"), accumulation and epilogues are fp32;
 *   - there is no CPU fallback: without a CUDA device every launch returns AF_ERR_CUDA.
 */
#ifndef ADAFOCUS_B200_H_
#define ADAFOCUS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AF_VERSION 210

typedef enum af_status {
  AF_OK = 0,
  AF_ERR_INVALID = -1, /* bad argument / unsupported shape */
  AF_ERR_CUDA = -2,    /* CUDA runtime / driver error */
  AF_ERR_STATE = -3    /* call not valid in the current recording state */
} af_status;

typedef enum af_act { AF_ACT_NONE = 0, AF_ACT_RELU = 1, AF_ACT_RELU6 = 2 } af_act;

typedef struct af_ctx af_ctx;   /* one per device; calls on one ctx must be externally serialised */
typedef struct af_plan af_plan; /* a recorded sequence of launches, replayable with fixed pointers */

/* Convolution / GEMM descriptor for af_conv2d_nhwc_f16 (the tcgen05 implicit-GEMM kernel).
 *   in : NHWC fp16 (n, h, w, cin), pixel stride in_stride elements (>= cin, multiple of 8)
 *   w  : packed fp16 [cout_pad][kh*kw*cblk*64] with cblk = ceil(cin/64); k = ((r*kw+s)*cblk*64 + ci), zero padded;
 *        cout_pad = ceil(cout/block_n)*block_n
 *   out: NHWC fp16 or fp32 (n, ho, wo, cout), pixel stride out_stride
 *   y  = act(scale[co]*conv + bias[co] (+ residual)), scale/bias have cout_pad entries; scale may be NULL (= 1,
 *        e.g. when the BatchNorm scale has been folded into the packed weights).  With a residual, scale == NULL is
 *        the fast path (the residual tile is TMA-loaded and accumulated on the tensor core, exact in fp32); a
 *        residual together with an explicit scale falls back to a slower epilogue that adds it from registers.
 * A plain GEMM C[M,N] = A[M,K] W[N,K]^T is the case n=1, h=1, w=M, cin=K, kh=kw=1, stride=1, pad=0. */
typedef struct af_conv_desc {
  const void* in;
  const void* w;
  const float* scale;
  const float* bias;
  const void* residual; /* NHWC fp16 with pixel stride res_stride, or NULL */
  void* out;
  int32_t n, h, w_, cin;
  int32_t cout;
  int32_t kh, kw, stride, pad;
  int32_t block_n; /* multiple of 16, <= 256 */
  int32_t act;     /* af_act */
  int32_t out_f32; /* 0: fp16 output, 1: fp32 output */
  int64_t in_stride, out_stride, res_stride;
  /* Optional explicit row / image strides of `in` (elements; 0 = dense: w_*in_stride, h*w_*in_stride).  With them
   * set, in_stride may be smaller than cin: consecutive "pixels" then overlap in memory (a sliding window over a
   * narrower tensor, e.g. the space-to-depth stem input of af_stem_s2d).  stride-1 convolutions only. */
  int64_t in_row_stride, in_img_stride;
  /* Temporal shift (STH/ops/temporal_shift.py:29-46, TemporalShift.shift) folded into the input loads of a 1x1
   * stride-1 convolution: the n images are n / tsm_t clips of tsm_t frames; input channels [0, tsm_fold) are read
   * from the next frame, [tsm_fold, 2*tsm_fold) from the previous one, the rest from the frame itself, zeros outside
   * the clip -- conv(shift(x)) without materialising shift(x).  tsm_t = 0: off.  Needs af_conv_tsm_supported(). */
  int32_t tsm_t, tsm_fold;
  /* pool != 0: the stem of the focus network -- conv + BN + ReLU + MaxPool2d(kernel 3, stride 2, padding 1)
   * (ACT/models/resnet.py:138-142, 213-216) as ONE kernel: `out` is then the pooled tensor (n, Ho/2, Wo/2, cout) and the
   * convolution's own output never reaches HBM.  Needs a stride-1 filter with kh >= 2 and resident weights, Wo == 64,
   * Ho even, cout <= 64, a ReLU-type activation (the pool pads with 0). */
  int32_t pool, reserved0;
  /* Optional second GEMM accumulated into the same output tile before the epilogue -- the projection shortcut of a
   * bottleneck block, out = relu(bn3(conv3(h)) + bn_d(conv_d(x))) (ACT/models/resnet.py:94-114 with `downsample`,
   * :170-181): in2 = the block input x (n, h2, w2_, cin2) NHWC fp16 with pixel stride in2_stride, w2 = the 1x1
   * stride-`stride2` downsample weights packed like `w` ([cout_pad][ceil(cin2/64)*64], BN scale folded in), `bias` =
   * the sum of both folded BN biases; needs scale == NULL and residual == NULL.  The downsample output and its
   * read-back as a residual never touch HBM.  in2 == NULL: off. */
  const void* in2;
  const void* w2;
  int32_t cin2, stride2, h2, w2_;
  int64_t in2_stride;
} af_conv_desc;

/* 1 when af_conv2d_nhwc_f16 can fold the temporal shift for this geometry (cin % 64 == 0, fold % 16 == 0,
 * n % t == 0, dense input); otherwise use af_tsm_shift_nhwc_f16 in front of the convolution. */
int af_conv_tsm_supported(int n, int h, int w, int cin, int in_stride, int fold, int t);

int af_version(void);
const char* af_last_error(void);

int af_ctx_create(af_ctx** out, int device);
int af_ctx_destroy(af_ctx* ctx);
int af_ctx_sm_count(const af_ctx* ctx);

/* Recording: between af_plan_begin and af_plan_end every kernel entry point called on ctx is appended to the plan
 * (tensor maps and launch geometry are resolved once) instead of being launched.  af_plan_run replays the whole
 * sequence natively on a stream -- this is how the T-step policy loop and the ~200 layer launches of one forward
 * (ACT/models/gfv_net.py:95-133) run without returning to Python. */
int af_plan_begin(af_ctx* ctx);
int af_plan_end(af_ctx* ctx, af_plan** out);
int af_plan_run(af_plan* plan, void* stream); /* CUDA-graph capturable; timing marks are skipped under capture */
int af_plan_num_launches(const af_plan* plan); /* kernel launches per replay (marks excluded) */
/* Timing marks: while recording, af_plan_mark inserts a CUDA event record at this point of the sequence; after a
 * replay has completed, af_plan_mark_elapsed_ms gives the device time between two marks of that replay. */
int af_plan_mark(af_ctx* ctx, int* mark_index);
int af_plan_mark_elapsed_ms(af_plan* plan, int mark_a, int mark_b, float* ms);
int af_plan_destroy(af_plan* plan);

/* Whole-path forward through the C ABI (SURVEY.md section 8(b) af_gfv_forward / af_workspace_bytes).  The layer schedule
 * and the weight packing are host logic of the Python shim (north_star: "host code stays Python/PyTorch calling ...
 * through a thin C-ABI extension"); once the shim has recorded a plan for a batch size, the plan IS the native forward
 * of GFV.forward(one_step=True) (ACT/models/gfv_net.py:95-133): af_plan_bind_forward names the plan's static input /
 * scan buffers and its padded logits buffer (rows = B*T of row_stride floats, `classes` valid, `steps` = T), and
 * af_gfv_forward copies caller tensors in (device pointers; NULL or the bound buffer itself = already in place),
 * replays the ~170 launches on `stream`, and writes logits (B*T, C) and / or last_out (B, C) contiguously.
 * No host synchronisation, CUDA-graph capturable, no allocation. */
int af_plan_bind_forward(af_plan* plan, void* input_buf, size_t input_bytes, void* scan_buf, size_t scan_bytes,
                         const float* logits_buf, int rows, int row_stride, int classes, int steps,
                         size_t workspace_bytes);
size_t af_workspace_bytes(const af_plan* plan);
int af_gfv_forward(af_plan* plan, const float* input, const float* scan, float* logits, float* last_out, void* stream);

/* get_patch(images, action_sequence, patch_size) -- ACT/models/utils.py:37-51 (= STH/models/utils.py:44-58).
 * img (N,C,H,W) fp32 NCHW -> out (N,C,P,P) fp32.  Exactly one of action (N,2 fp32 in [0,1], coordinates computed
 * as floor(a*(H-P)) in fp32 like the reference) and yx (N,2 int32) is non-NULL.  yx_out (N,2 int32) optional. */
int af_crop_nchw_f32(af_ctx* ctx, const float* img, const float* action, const int32_t* yx, float* out,
                     int32_t* yx_out, int N, int C, int H, int W, int P, void* stream);

/* floor(action*(H-P)).int() -- ACT/models/utils.py:42. */
int af_action_to_yx(af_ctx* ctx, const float* action, int32_t* yx, int N, int H, int P, void* stream);

/* Crop + fp32->fp16 + im2col staging that feeds a 3-channel stem convolution as a GEMM:
 * ResNet conv1 7x7/2 pad 3 (ACT/models/resnet.py:138) on the patch selected by yx (fused get_patch), or
 * MobileNet-V2 features[0] 3x3/2 pad 1 (ACT/models/mobilenet.py:105) on the whole frame (yx = NULL, P = H).
 * frames (N,3,H,W) fp32 -> out [(N*Ho*Wo)][Kpad] fp16, k = (r*KW+s)*3 + c.  yx has one (y,x) row per yx_div
 * consecutive frames (1 for ACT; T_f / video_div for STH, where get_patch crops all frames of a division at once,
 * STH/models/gfv_net.py:141-152,421). */
int af_stem_im2col(af_ctx* ctx, const float* frames, const int32_t* yx, int yx_div, void* out, int N, int H, int W,
                   int P, int KH, int KW, int stride, int pad, int Kpad, void* stream);

/* get_patch + conv zero padding + 2x2 space-to-depth + fp32->fp16 for a stride-2 3-channel stem convolution
 * (ResNet conv1 7x7/2 pad 3, ACT/models/resnet.py:138; MobileNet-V2 features[0] 3x3/2 pad 1,
 * ACT/models/mobilenet.py:105).  frames (N,3,H,W) fp32 -> out (N,Hs,Ws,16) fp16 with
 * out[n][Y][X][(dy*2+dx)*3+c] = padded_patch[c][2Y+dy][2X+dx] (zero outside the P x P patch, channels 12-15 zero).
 * The stride-2 KxK conv is then a stride-1 ceil(K/2) x 1 af_conv2d_nhwc_f16 over the 64-channel sliding-window view
 * (in_stride = 16, in_row_stride = Ws*16) of this tensor.  vt = 2 also folds the vertical neighbour into the pixel
 * (out (N,Hs,Ws,32), channel v*16 + (dy*2+dx)*3 + c = padded[c][2(Y+v)+dy][2X+dx]): a 3x3/2 stem is then a 1x1 conv
 * over the two-position window view (in_stride = 32).  yx / yx_div as for af_stem_im2col. */
int af_stem_s2d(af_ctx* ctx, const float* frames, const int32_t* yx, int yx_div, void* out, int N, int H, int W, int P,
                int pad, int Hs, int Ws, int vt, void* stream);

/* Fused get_patch + stem convolution + BN + activation as ONE tcgen05 implicit-GEMM kernel: the patch selected by yx
 * (ACT/models/utils.py:37-51) goes through Conv2d(3, cout, KHxKW, stride, pad) (ResNet conv1, ACT/models/resnet.py:138)
 * without the im2col matrix ever reaching HBM -- producer warps build the swizzled A tiles in shared memory from the
 * fp32 NCHW frames.  w: packed fp16 [cout][ceil(KH*KW*3/64)*64] with k = (r*KW+s)*3 + c (as for af_stem_im2col);
 * out: NHWC fp16 (N, Ho, Wo, cout).  Needs cout % 16 == 0, cout <= 64, KH*KW*3 <= 256, Ho*Wo >= 128. */
int af_stem_conv_fused(af_ctx* ctx, const float* frames, const int32_t* yx, int yx_div, const void* w,
                       const float* scale, const float* bias, void* out, int N, int H, int W, int P, int cout, int KH,
                       int KW, int stride, int pad, int act, void* stream);

/* Fused MobileNet-V2 inverted-residual block (ACT/models/mobilenet.py:42-68, InvertedResidual.forward with
 * expand_ratio != 1): 1x1 expand + BN + ReLU6 -> depthwise 3x3 (stride 1|2, pad 1) + BN + ReLU6 -> 1x1 project + BN
 * (+ residual) as ONE kernel; the expanded tensor stays in shared memory / TMEM.
 *   in  : NHWC fp16 (n, h, w, cin) contiguous            out : NHWC fp16 (n, ho, wo, cout) contiguous
 *   w1  : packed fp16 [ceil(cexp/64)*64][64] (k = ci, BN scale folded in, zero padded)   bias1: fp32 [ceil(cexp/64)*64]
 *   dw_w: fp32 [9][ceil(cexp/64)*64] with the BN scale folded in (tap-major), zero padded; bias2 likewise
 *   w2  : packed fp16 [ceil(cout/16)*16][ceil(cexp/64)*64] (BN scale folded in)           bias3: fp32 [ceil(cout/16)*16]
 *   residual: NHWC fp16 (n, ho, wo, cout) with pixel stride res_stride, or NULL.
 * af_mbconv_fused_supported is a host-side query (no device work): 1 when the shape is handled (cin, cout <= 64 and
 * multiples of 8; cexp a multiple of 16 whose last 64-channel chunk holds 16, 32 or 64 channels), else 0. */
typedef struct af_mbconv_desc {
  const void* in;
  const void* w1;
  const float* bias1;
  const float* dw_w;
  const float* bias2;
  const void* w2;
  const float* bias3;
  const void* residual;
  void* out;
  int32_t n, h, w_, cin, cexp, cout, stride;
  int64_t res_stride;
  /* 1: the expand bias rides in the weights -- w1[:, cin] = fp16(bias1), w1[:, cin + 1] = fp16(bias1 - fp16(bias1)),
   * needs cin + 2 <= 64; the kernel then plants a 1.0 pair in those two channels of every in-image pixel of the input
   * tile, so the bias (to ~22 bits) and the zero padding of the expanded tensor both come out of the MMA and the
   * expand epilogue is a pure ReLU6 + convert.  bias1 is ignored. */
  int32_t bias1_in_w1;
} af_mbconv_desc;
int af_mbconv_fused_supported(int n, int h, int w, int cin, int cexp, int cout, int stride);
int af_mbconv_fused(af_ctx* ctx, const af_mbconv_desc* d, void* stream);

/* Row-streaming form of the same block for large batches (csrc/mbconv_rows.cuh): the expand GEMM runs transposed
 * (TMEM lane = expanded channel, column = pixel), every thread of the depthwise warps owns one expanded channel, reads
 * its image rows straight out of TMEM and rolls them through registers while the CTA walks down whole frames; the
 * 6x expanded tensor touches neither HBM nor shared memory.  Shapes: cin, cout <= 64; stride 1 with w in {14, 28, 56};
 * (and 112 as two 56-pixel segments per row); stride 2 with w in {28, 56, 112}; cexp % 16 == 0 with at most three 128-lane chunks (af_mbconv_rows_supported).
 * The placement of expanded channels on TMEM lanes depends on cexp and on the number of 14-output column strips per
 * row (spr = 1, 2 or 4: stride 1 -> w / 14, stride 2 -> min(w / 14, 4)); af_mbconv_rows_layout returns it so that the
 * host packs
 *   w1  : fp16 [nchunks*128][64]   row (chunk, lane) = W1[lane_ch] * bn_scale / 6, zero rows for lane_ch < 0
 *   dwp : fp32 [nchunks][11][128]  per lane: 9 depthwise taps (BN scale folded), bias1 / 6, bias2 / 6
 *   w2  : fp16 [ceil(cout/16)*16][nchunks*128]   column chunk*128 + lane_kpos = 6 * W2[:, lane_ch] * bn_scale
 *   bias3: fp32 [ceil(cout/16)*16]
 * (ReLU6 is evaluated as 6 * saturate(x / 6)).  lane_ch / lane_kpos: int16 [3][128]. */
typedef struct af_mbconv_rows_desc {
  const void* in;
  const void* w1;
  const float* dwp;
  const void* w2;
  const float* bias3;
  const void* residual;
  void* out;
  int32_t n, h, w_, cin, cexp, cout, stride;
  int64_t res_stride;
  /* input view, in elements (0 = contiguous NHWC): the "pixel" of the view may be wider than its stride (overlapping
   * pixels), as in af_conv_desc.in_row_stride -- the stem of MobileNet-V2 (stride-2 3x3 conv = 1x1 conv over the
   * 64-channel window view of the space-to-depth image af_stem_s2d writes) then runs as the expand GEMM of a block:
   * stem conv + BN + ReLU6 -> depthwise 3x3 + BN + ReLU6 -> 1x1 project + BN of features[0..1] in one launch. */
  int64_t in_pix_stride, in_row_stride, in_img_stride;
} af_mbconv_rows_desc;
int af_mbconv_rows_supported(int n, int h, int w, int cin, int cexp, int cout, int stride);
int af_mbconv_rows_layout(int cexp, int spr, int32_t* nchunks, int16_t* lane_ch, int16_t* lane_kpos);
int af_mbconv_rows(af_ctx* ctx, const af_mbconv_rows_desc* d, void* stream);
/* bring-up hook: device buffer (32 x 8 int64) for per-warp cycle counters of CTA 0 of af_mbconv_rows (only filled by
 * instrumented builds of csrc/mbconv_rows.cu; the shipped kernel ignores it), or NULL */
int af_debug_mbconv_rows_prof(void* buf);

/* MobileNet-V2 features[0]: Conv2d(3,32,3,stride 2,pad 1) + BN + ReLU6 (ACT/models/mobilenet.py:105) directly from
 * the fp32 NCHW frames to NHWC fp16 on the FMA pipes (K = 27 is too thin for a tensor-core k-block).
 * w27 fp32 [27][32] with k = (r*3+s)*3 + c; scale / bias fp32 [32]. */
int af_stem_conv3x3s2_c32(af_ctx* ctx, const float* frames, const float* w27, const float* scale, const float* bias,
                          void* out, int N, int H, int W, int act, void* stream);

/* Conv2d + BatchNorm2d(eval, folded) + ReLU/ReLU6 (+ residual add) -- ACT/models/resnet.py:94-114,
 * ACT/models/mobilenet.py:32-68, ACT/models/ppo.py:33-39; also every Linear/GRU projection as a 1x1 "conv". */
int af_conv2d_nhwc_f16(af_ctx* ctx, const af_conv_desc* desc, void* stream);

/* Depthwise 3x3 pad 1 + folded BN + activation (ACT/models/mobilenet.py:58); w9c fp32 [9][C]. */
int af_dwconv3x3_nhwc_f16(af_ctx* ctx, const void* in, const float* w9c, const float* scale, const float* bias,
                          void* out, int N, int H, int W, int C, int stride, int act, void* stream);

/* MaxPool2d(3,2,1) -- ACT/models/resnet.py:141. */
int af_maxpool3x3s2_nhwc_f16(af_ctx* ctx, const void* in, void* out, int N, int H, int W, int C, void* stream);

/* Global average pool -- ACT/models/resnet.py:223 (avgpool), ACT/models/mobilenet.py:148 (mean([2,3])).
 * Writes fp32 rows (stride out_f32_stride) and/or fp16 rows (stride out_f16_stride); this is also how the
 * [global | local] feature concat of ACT/models/gfv_net.py:121,132 is assembled without torch.cat. */
int af_avgpool_nhwc_f16(af_ctx* ctx, const void* in, float* out_f32, int64_t out_f32_stride, void* out_f16,
                        int64_t out_f16_stride, int N, int HW, int C, void* stream);

int af_nhwc_f16_to_nchw_f32(af_ctx* ctx, const void* in, float* out, int N, int HW, int C, void* stream);
int af_nchw_f32_to_nhwc_f16(af_ctx* ctx, const float* in, void* out, int N, int C, int HW, int Cpad, void* stream);

/* One GRU time step's gate math (torch.nn.GRU, gates r,z,n) -- ACT/models/ppo.py:80, ACT/models/gfv_net.py:431.
 * xg (B,3H) row stride xg_stride = W_ih x + b_ih; hg (B,3H) = W_hh h + b_hh (both produced by af_conv2d_nhwc_f16). */
int af_gru_gates(af_ctx* ctx, const float* xg, int64_t xg_stride, const float* hg, const float* h_prev, float* h_new,
                 void* h_new_f16, void* hseq_f16, int64_t hseq_stride, float* hseq_f32, int64_t hseq_f32_stride,
                 int B, int Hd, int split, void* stream);

/* fp32 rows (row stride in_stride) -> split-precision fp16 operand rows [hi | lo | hi], 3*cols wide, x = hi + lo.
 * Multiplied with weights packed [W_hi | W_hi | W_lo] (engine.pack_conv_split) the tensor core accumulates
 * x_hi W_hi + x_lo W_hi + x_hi W_lo in fp32: ~22-bit operands for the classifier head (ACT/models/gfv_net.py:427-435),
 * which is < 1 % of the FLOPs but sets the logit error.  `split` != 0 in af_gru_gates / af_gru_sequence makes their
 * fp16 outputs (h_new_f16 row stride 3*Hd, hseq_f16) use the same three-part rows and, for af_gru_sequence, reads
 * w_hh_f16 as (3H, 3H) [W_hi | W_hi | W_lo] rows. */
int af_split3_f16(af_ctx* ctx, const float* in, int64_t in_stride, void* out, int rows, int cols, void* stream);

/* A whole GRU sequence (all T steps, h0 = zeros when NULL) in ONE persistent cooperative launch, for small batches:
 * the policy rollout of ACT/models/ppo.py:67-96 / ACT/models/gfv_net.py:110 and the classifier GRU of
 * ACT/models/gfv_net.py:427-435 without leaving the device between steps.  xg (B*T,3H) fp32 rows b*T+t = W_ih x + b_ih;
 * w_hh_f16 (3H,H) fp16 row-major; hbuf: 2*B*H floats scratch; counter: one uint32 scratch.  Outputs h_t as fp16 rows
 * b*T+t of hseq_f16 (row stride hseq_stride) and optionally the final fp32 state h_out (B,H).
 * Constraints: H % 256 == 0, H <= 1024 and H / 8 <= number of SMs. */
int af_gru_sequence(af_ctx* ctx, const float* xg, const void* w_hh_f16, const float* b_hh, const float* h0, float* hbuf,
                    void* hseq_f16, int64_t hseq_stride, float* h_out, uint32_t* counter, int B, int T, int Hd,
                    int split, void* stream);

/* The same recurrence for up to 64 sequences on the tensor core, still ONE persistent cooperative launch (csrc/gru_tc.cuh):
 * CTA c keeps the 24 W_hh rows of hidden units [8c, 8c+8) in shared memory as UMMA operand tiles, streams h_{t-1}
 * (fp16 operand rows, TMA) every step, accumulates in TMEM, applies the gates from fp32 state held in registers and
 * meets the other CTAs at a grid barrier.  w_hh_f16: (3H, w_cols) fp16 row-major -- plain W_hh (w_cols >= H) or the
 * split-precision packing [W_hi | W_hi | W_lo] (split != 0, w_cols >= 3H), in which case h travels as [hi | lo] and
 * hseq_f16 rows are [hi | lo | hi] (3H wide).  hbuf: 2 * B * (split ? 2 : 1) * H halfs of scratch; counter: one uint32.
 * Constraints (af_gru_sequence_tc_supported): 1 <= B <= 64, H % 64 == 0, H / 8 <= number of SMs. */
int af_gru_sequence_tc_supported(const af_ctx* ctx, int B, int Hd, int split);
int af_gru_sequence_tc(af_ctx* ctx, const float* xg, const void* w_hh_f16, int64_t w_cols, const float* b_hh,
                       const float* h0, void* hbuf, void* hseq_f16, int64_t hseq_stride, float* h_out, uint32_t* counter,
                       int B, int T, int Hd, int split, void* stream);

/* softmax -> argmax -> standard action table -> patch origin; ACT/models/ppo.py:84,94,
 * ACT/models/gfv_net.py:272-307,345-347, ACT/models/utils.py:42.  grid_n = sqrt(action_dim). */
int af_policy_head(af_ctx* ctx, const float* logits, int64_t logit_stride, int A, int grid_n, int rows, int H, int P,
                   int32_t* action_idx, float* action_yx, int32_t* yx, void* stream);

/* STH continuous policy head: action_mean = sigmoid(.) -- STH/models/ppo_continuous.py:61-63,106-107. */
int af_policy_head_continuous(af_ctx* ctx, const float* logits, int64_t logit_stride, int rows, int H, int P,
                              float* action_yx, int32_t* yx, void* stream);

/* TemporalShift.shift -- STH/ops/temporal_shift.py:29-46, NHWC fp16, fold = C / shift_div. */
int af_tsm_shift_nhwc_f16(af_ctx* ctx, const void* in, void* out, int NT, int T, int HW, int C, int fold,
                          void* stream);

/* The same shift on reference-layout tensors (NT,C,H,W) fp32 -- the public TemporalShift.shift(). */
int af_tsm_shift_nchw_f32(af_ctx* ctx, const float* in, float* out, int NT, int T, int C, int HW, int fold,
                          void* stream);

/* ConsensusModule('avg') (+ glancer consensus) -- STH/ops/basic_ops.py:18-27, STH/models/gfv_net.py:170-172. */
int af_consensus_avg(af_ctx* ctx, const float* in, const float* add, float* out, int B, int T, int C, void* stream);

/* Evaluation metrics on the device, consumers of the path's logits.
 * af_topk_hits: hits[0] += #rows whose target ranks < k0, hits[1] += ... < k1 -- accuracy(), ACT/ops/utils.py:35-49.
 * af_softmax_rows + af_class_ap: per-class average precision over N samples with (N, L) int64 labels (-1 = none) --
 * cal_map(), ACT/ops/utils.py:68-88 (ranks by counting instead of sorting; ties: lower index first). */
int af_topk_hits(af_ctx* ctx, const float* logits, int64_t stride, const int64_t* target, int rows, int C, int k0,
                 int k1, float* hits, void* stream);
int af_softmax_rows(af_ctx* ctx, const float* logits, int64_t stride, float* probs, int rows, int C, void* stream);
int af_class_ap(af_ctx* ctx, const float* probs, const int64_t* labels, int N, int C, int L, float* ap, void* stream);

int af_fill_f32(af_ctx* ctx, float* p, float v, int64_t n, void* stream);
int af_f32_to_f16(af_ctx* ctx, const float* in, void* out, int64_t n, void* stream);

/* Frame ingest -- Stack -> ToTorchFormatTensor(div=True) -> GroupNormalize of the reference's loaders
 * (ACT/ops/transforms.py:303-336, 64-77) on the device, so that clips cross PCIe as bytes:
 * in (B, HW, C) uint8 with C = 3T channels in frame-major RGB order (what np.concatenate(frames, axis=2) produces)
 * -> out (B, C, HW) fp32 = ((u / 255) - mean[c % 3]) / std[c % 3], each step rounded to fp32 (bit-identical to the
 * torch ops).  mean3 / std3 are HOST pointers to three floats. */
int af_frames_u8_to_f32(af_ctx* ctx, const uint8_t* in, float* out, int B, int HW, int C, const float* mean3,
                        const float* std3, void* stream);

/* GroupScale -> GroupCenterCrop (-> Stack) of the reference's validation transform (ACT/ops/transforms.py:78-93, 37-43,
 * 303-316; ACT/main_dist.py:213-217) on decoded uint8 frames, on the device and bit-identical to Pillow's
 * Image.resize(BILINEAR) behind torchvision.transforms.Resize: two integer passes (horizontal into `tmp`, then
 * vertical) with the window bounds and 22-bit fixed-point weights Pillow's precompute_coeffs / normalize_coeffs_8bpc
 * produce -- the caller computes them on the host (adafocus_b200.preprocess.resize_tables) for the OW columns / OH rows
 * the centre crop keeps.
 *   in   (N, H, W, C) uint8 frames (W * C <= 12288);  tmp (N, rows, OW, C) scratch for source rows [row0, row0 + rows);
 *   out  (N, OH, OW, C) uint8.  af_frames_u8_to_f32 over these N frames (B = N, C channels) then writes (N, C, OH*OW)
 *        fp32, which is the (clips, T*C, OH, OW) tensor Stack() + ToTorchFormatTensor produce for T frames per clip.
 *   hbounds / vbounds int32 [OW|OH][2] = (first source index, taps), hkk / vkk int32 [OW|OH][hks|vks]: DEVICE pointers. */
int af_resize_crop_u8(af_ctx* ctx, const uint8_t* in, uint8_t* tmp, uint8_t* out, int N, int H, int W, int C,
                      const int32_t* hbounds, const int32_t* hkk, int hks, int OW, const int32_t* vbounds,
                      const int32_t* vkk, int vks, int OH, int row0, int rows, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADAFOCUS_B200_H_ */
