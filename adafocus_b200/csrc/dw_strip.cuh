// Register-blocked depthwise 3x3 inner loop shared by the stand-alone depthwise kernel (dwconv_tma.cu) and the fused
// inverted-residual block kernel (mbconv_fused.cu).  Device-only, header-only.
#pragma once
#include <cuda_fp16.h>
#include <cstdint>

namespace af {

// packed fp32 FMA (Blackwell FFMA2): two channels per instruction
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t"
      ".reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "mov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t"
      "}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

// fp32 pair -> packed fp16 with the ReLU folded into the conversion (cvt.rn.relu.f16x2.f32: negative results become
// +0), one instruction instead of a conversion and a max.  The first PTX source operand is the UPPER half.
__device__ __forceinline__ __half2 floats2half2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return *reinterpret_cast<__half2*>(&d);
}

// One thread's share of a tile: 4 channels x 2 output columns x RO output rows.
//   in  : top-left input pixel of the thread's window (its 4 channels), pixel pitch pix_b bytes, row pitch row_b bytes
//   out0 / out1 : where the thread's left / right output pixel of row 0 goes (row pitch orow_b); two pointers so that
//                 the caller may use a swizzled destination layout
template <int S, int RO>
__device__ __forceinline__ void dw_strip(const uint8_t* __restrict__ in, int pix_b, int row_b, uint8_t* __restrict__ out0,
                                         uint8_t* __restrict__ out1, int orow_b, const float2 (&w)[9][2], const float2 (&bias)[2],
                                         int act) {
  constexpr int NCOLS = S + 3;            // input columns feeding two adjacent outputs
  constexpr int IN_ROWS = (RO - 1) * S + 3;
  float2 acc[RO][2][2];
#pragma unroll
  for (int r = 0; r < IN_ROWS; ++r) {
    float2 x[NCOLS][2];
#pragma unroll
    for (int cidx = 0; cidx < NCOLS; ++cidx) {
      const uint2 v = *reinterpret_cast<const uint2*>(in + r * row_b + cidx * pix_b);
      x[cidx][0] = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
      x[cidx][1] = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    }
#pragma unroll
    for (int kh = 2; kh >= 0; --kh) {
      if ((r - kh) < 0 || (r - kh) % S != 0) continue;
      const int o = (r - kh) / S;
      if (o >= RO) continue;
#pragma unroll
      for (int px = 0; px < 2; ++px) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float2 base = (kh == 0 && kw == 0) ? bias[h] : acc[o][px][h];
            acc[o][px][h] = ffma2(x[px * S + kw][h], w[kh * 3 + kw][h], base);
          }
        }
      }
      if (kh == 2) {
        // output row o is complete
#pragma unroll
        for (int px = 0; px < 2; ++px) {
          __half2 h0, h1;
          if (act != 0) {
            h0 = floats2half2_relu(acc[o][px][0].x, acc[o][px][0].y);
            h1 = floats2half2_relu(acc[o][px][1].x, acc[o][px][1].y);
            if (act == 2) {
              const __half2 six = __float2half2_rn(6.f);
              h0 = __hmin2(h0, six);
              h1 = __hmin2(h1, six);
            }
          } else {
            h0 = __floats2half2_rn(acc[o][px][0].x, acc[o][px][0].y);
            h1 = __floats2half2_rn(acc[o][px][1].x, acc[o][px][1].y);
          }
          uint2 ov;
          ov.x = *reinterpret_cast<const uint32_t*>(&h0);
          ov.y = *reinterpret_cast<const uint32_t*>(&h1);
          *reinterpret_cast<uint2*>((px == 0 ? out0 : out1) + o * orow_b) = ov;
        }
      }
    }
  }
}

}  // namespace af
