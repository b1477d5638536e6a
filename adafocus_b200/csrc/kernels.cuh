// HBM-bound helper kernels of the AdaFocus inference path (crop, stem staging, depthwise conv, pooling,
// GRU gates, policy head). Launchers only; see kernels.cu for the reference call sites.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace af {

// get_patch (ACT/models/utils.py:37-51): out[n] = img[n, :, y:y+P, x:x+P], (y,x) = floor(action*(H-P)) or given.
cudaError_t launch_crop_nchw_f32(const float* img, const float* action, const int32_t* yx, float* out,
                                 int32_t* yx_out, int N, int C, int H, int W, int P, cudaStream_t s);

// action (N,2) fp32 in [0,1] -> int32 (y,x) = floor(a * (H - P)) evaluated in fp32 like the reference.
cudaError_t launch_action_to_yx(const float* action, int32_t* yx, int N, int H, int P, cudaStream_t s);

// Crop + fp32->fp16 + im2col staging of a 3-channel NCHW frame for a KHxKW / stride / pad stem convolution.
// out[(n*Ho+oh)*Wo+ow][k], k = (kh*KW+kw)*3 + c (zero for k >= KH*KW*3 and for taps outside the P x P window).
// yx holds one (y,x) per yx_div consecutive frames.
cudaError_t launch_u8hwc_to_f32chw_norm(const uint8_t* in, float* out, int B, int HW, int C, const float* mean3,
                                        const float* std3, cudaStream_t s);
cudaError_t launch_stem_s2d(const float* frames, const int32_t* yx, int yx_div, __half* out, int N, int H, int W, int P,
                            int pad, int Hs, int Ws, int vt, cudaStream_t s);
cudaError_t launch_stem_im2col(const float* frames, const int32_t* yx, int yx_div, __half* out, int N, int H, int W,
                               int P, int KH, int KW, int stride, int pad, int Ho, int Wo, int Kpad, cudaStream_t s);

// Direct 3 -> 32 channel 3x3 / stride 2 / pad 1 conv + folded BN + activation, fp32 NCHW frames -> NHWC fp16
// (MobileNet-V2 features[0], ACT/models/mobilenet.py:105). w27: fp32 [27][32], k = (r*3+s)*3 + c.
cudaError_t launch_stem_conv3x3s2(const float* frames, const float* w27, const float* scale, const float* bias,
                                  __half* out, int N, int H, int W, int act, cudaStream_t s);

// Depthwise 3x3 (pad 1) + folded BN + ReLU6, NHWC fp16 (ACT/models/mobilenet.py:58, groups=hidden_dim).
cudaError_t launch_dwconv3x3(const __half* in, const float* w9c, const float* scale, const float* bias, __half* out,
                             int N, int H, int W, int C, int stride, int act, cudaStream_t s);

// MaxPool2d(3, stride 2, pad 1), NHWC fp16 (ACT/models/resnet.py:141).
cudaError_t launch_maxpool3x3s2(const __half* in, __half* out, int N, int H, int W, int C, cudaStream_t s);

// Global average pool over HW, NHWC fp16 -> fp32 and/or fp16 rows with arbitrary row stride
// (AdaptiveAvgPool2d((1,1)) ACT/models/resnet.py:223; x.mean([2,3]) ACT/models/mobilenet.py:148).
cudaError_t launch_avgpool(const __half* in, float* out_f32, long long out_f32_stride, __half* out_f16,
                           long long out_f16_stride, int N, int HW, int C, cudaStream_t s);

// NHWC fp16 -> NCHW fp32 (to hand glance() feature maps back in the reference's layout).
cudaError_t launch_nhwc_f16_to_nchw_f32(const __half* in, float* out, int N, int HW, int C, cudaStream_t s);

// NCHW fp32 -> NHWC fp16 with channel padding (generic entry for callers that bring their own patches).
cudaError_t launch_nchw_f32_to_nhwc_f16(const float* in, __half* out, int N, int C, int HW, int Cpad,
                                        cudaStream_t s);

// GRU gate math for one time step (torch.nn.GRU semantics, gate order r,z,n):
//   r = sig(xr+hr) z = sig(xz+hz) n = tanh(xn + r*hn) h' = (1-z)*n + z*h
// xg: [B,3H] row stride xg_stride (includes b_ih), hg: [B,3H] contiguous (includes b_hh).
cudaError_t launch_gru_gates(const float* xg, long long xg_stride, const float* hg, const float* h_prev,
                             float* h_new, __half* h_new_f16, __half* hseq_f16, long long hseq_stride,
                             float* hseq_f32, long long hseq_f32_stride, int B, int Hd, int split, cudaStream_t s);
cudaError_t launch_split3_f16(const float* in, long long in_stride, __half* out, int rows, int cols, cudaStream_t s);

// Whole GRU sequence in one persistent launch (small batches): xg [B*T,3H] fp32 rows b*T+t (W_ih x + b_ih), w_hh fp16
// [3H][H], b_hh fp32 [3H], h0 [B,H] or null (zeros); hbuf = 2*B*H floats of scratch, counter = one zero-initialised
// uint32 (reset by the launcher).  Writes h_t as fp16 rows b*T+t of hseq_f16 and the final state to h_out (optional).
cudaError_t launch_gru_sequence(const float* xg, const __half* w_hh, const float* b_hh, const float* h0, float* hbuf,
                                __half* hseq_f16, long long hseq_stride, float* h_out, unsigned int* counter, int B,
                                int T, int Hd, int sm_count, int split, cudaStream_t s);

// softmax over A logits, argmax (first maximum), action table lookup, floor(a*(H-P)) -> int32 (y,x).
// ACT/models/ppo.py:84,94 + ACT/models/gfv_net.py:345-347 + ACT/models/utils.py:42.
cudaError_t launch_policy_head(const float* logits, long long logit_stride, int A, int grid_n, int rows,
                               int H, int P, int32_t* action_idx, float* action_yx, int32_t* yx, cudaStream_t s);

// STH continuous policy head: sigmoid(actor logits) (STH/models/ppo_continuous.py:61-63,106-107) -> (y,x).
cudaError_t launch_policy_head_continuous(const float* logits, long long logit_stride, int rows, int H, int P,
                                          float* action_yx, int32_t* yx, cudaStream_t s);

// Temporal shift (STH/ops/temporal_shift.py:29-46) on NHWC fp16: out[n,t,:,:,c] = in[n,t+1] for c<fold,
// in[n,t-1] for fold<=c<2fold, in[n,t] otherwise; zero at the clip ends.
cudaError_t launch_tsm_shift(const __half* in, __half* out, int NT, int T, int HW, int C, int fold,
                             cudaStream_t s);

cudaError_t launch_tsm_shift_nchw_f32(const float* in, float* out, int NT, int T, int C, int HW, int fold,
                                      cudaStream_t s);

// out[b, c] = mean_t in[b*T+t, c] (+ add[b, c]) ; STH ConsensusModule('avg') (STH/ops/basic_ops.py:18-27).
cudaError_t launch_consensus_avg(const float* in, const float* add, float* out, int B, int T, int C,
                                 cudaStream_t s);

// Evaluation metrics on the device: top-k hit counts (ACT/ops/utils.py:35-49), row softmax and
// per-class average precision (ACT/ops/utils.py:68-88).
cudaError_t launch_topk_hits(const float* logits, long long stride, const long long* target, int rows, int C, int k0,
                             int k1, float* hits, cudaStream_t s);
cudaError_t launch_softmax_rows(const float* logits, long long stride, float* probs, int rows, int C, cudaStream_t s);
cudaError_t launch_class_ap(const float* probs, const long long* labels, int N, int C, int L, float* ap,
                            cudaStream_t s);

cudaError_t launch_fill_f32(float* p, float v, long long n, cudaStream_t s);
cudaError_t launch_f32_to_f16(const float* in, __half* out, long long n, cudaStream_t s);
// GroupScale + GroupCenterCrop (+ Stack) on decoded uint8 frames, bit-identical to Pillow's bilinear resize
cudaError_t launch_pil_resize_crop_u8(const uint8_t* in, uint8_t* tmp, uint8_t* out, int N, int H, int W, int C,
                                      const int32_t* hbounds, const int32_t* hkk, int hks, int OW,
                                      const int32_t* vbounds, const int32_t* vkk, int vks, int OH, int row0, int rows,
                                      cudaStream_t s);

}  // namespace af
