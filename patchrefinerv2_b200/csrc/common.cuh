// Shared helpers for the sm_100a kernels: error plumbing, bf16 (hi,lo) activation access.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/prv2_b200.h"

namespace prv2 {

void set_error(const char* fmt, ...);

#define PRV2_CHECK_ARG(cond, ...)                     \
  do {                                                \
    if (!(cond)) {                                    \
      prv2::set_error(__VA_ARGS__);                   \
      return PRV2_EINVAL;                             \
    }                                                 \
  } while (0)

#define PRV2_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      prv2::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return PRV2_ECUDA;                                                             \
    }                                                                                \
  } while (0)

#define PRV2_LAUNCH_CHECK() PRV2_CUDA(cudaGetLastError())

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16 f2bf(float v) { return __float2bfloat16_rn(v); }

// Activation element access.  Two storage formats share the 16-bit planes:
//   lo == nullptr : ONE plane of bf16 (one-pass mode);
//   lo != nullptr : a (hi, lo) pair of FP16 planes, value = hi + lo (fp32-class mode).  fp16 pairs carry ~22 mantissa bits
//                   (bf16 pairs: 16) for |x| >= 2^-3; below that the lo plane goes subnormal and the ABSOLUTE error floors at
//                   3e-8, which is what matters for operands of dot products.  Stores saturate at +-65504 instead of overflowing.
#include <cuda_fp16.h>
constexpr float F16_MAX = 65504.0f;
__device__ __forceinline__ float sat16(float v) { return fminf(fmaxf(v, -F16_MAX), F16_MAX); }
__device__ __forceinline__ float act_load(const bf16* hi, const bf16* lo, size_t i) {
  if (lo) return __half2float(reinterpret_cast<const __half*>(hi)[i]) + __half2float(reinterpret_cast<const __half*>(lo)[i]);
  return bf2f(hi[i]);
}
__device__ __forceinline__ void act_store(bf16* hi, bf16* lo, size_t i, float v) {
  if (lo) {
    v = sat16(v);
    const __half h = __float2half_rn(v);
    reinterpret_cast<__half*>(hi)[i] = h;
    reinterpret_cast<__half*>(lo)[i] = __float2half_rn(v - __half2float(h));
  } else {
    hi[i] = f2bf(v);
  }
}

// 8-wide (16 byte) vector access
struct alignas(16) bf16x8 { bf16 v[8]; };

// 8 packed 16-bit values already in registers -> floats (f16 = the plane pair format, else bf16)
__device__ __forceinline__ void unpack8(const uint4& a, bool f16, float (&out)[8], bool add) {
  float t[8];
  if (f16) {
    const __half2* a2 = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(a2[k]); t[2 * k] = f.x; t[2 * k + 1] = f.y; }
  } else {
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(a2[k]); t[2 * k] = f.x; t[2 * k + 1] = f.y; }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) out[k] = add ? out[k] + t[k] : t[k];
}
__device__ __forceinline__ void act_load8(const bf16* hi, const bf16* lo, size_t i, float (&out)[8]) {
  const uint4 a = *reinterpret_cast<const uint4*>(hi + i);
  unpack8(a, lo != nullptr, out, false);
  if (lo) {
    const uint4 b = *reinterpret_cast<const uint4*>(lo + i);
    unpack8(b, true, out, true);
  }
}
__device__ __forceinline__ void act_store8(bf16* hi, bf16* lo, size_t i, const float (&in)[8]) {
  if (lo) {
    __half2 h2[4], l2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = sat16(in[2 * k]), b = sat16(in[2 * k + 1]);
      h2[k] = __floats2half2_rn(a, b);
      const float2 f = __half22float2(h2[k]);
      l2[k] = __floats2half2_rn(a - f.x, b - f.y);
    }
    *reinterpret_cast<uint4*>(hi + i) = *reinterpret_cast<const uint4*>(h2);
    *reinterpret_cast<uint4*>(lo + i) = *reinterpret_cast<const uint4*>(l2);
  } else {
    // packed conversions (F2FP.BF16.PACK_AB: two values per instruction)
    __nv_bfloat162 h2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) h2[k] = __floats2bfloat162_rn(in[2 * k], in[2 * k + 1]);
    *reinterpret_cast<uint4*>(hi + i) = *reinterpret_cast<const uint4*>(h2);
  }
}

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// ATen upsample_bilinear2d(align_corners=True) source coordinates, fp32 (UpSample.h).
struct BilinearTap { int i0, i1; float l0, l1; };
__device__ __forceinline__ float ac_scale(int n_in, int n_out) {
  return n_out > 1 ? __fdiv_rn((float)(n_in - 1), (float)(n_out - 1)) : 0.0f;
}
__device__ __forceinline__ BilinearTap ac_tap(float scale, int dst, int n_in) {
  BilinearTap t;
  float src = __fmul_rn(scale, (float)dst);
  t.i0 = min((int)src, n_in - 1);
  t.i1 = t.i0 + (t.i0 < n_in - 1 ? 1 : 0);
  t.l1 = __fsub_rn(src, (float)t.i0);
  t.l0 = __fsub_rn(1.0f, t.l1);
  return t;
}
// ATen CPU ordering (verified bit-exact, oracle/pr_oracle.py np_bilinear_ac):
// r0=fma(lx0,a,lx1*b); r1=fma(lx0,c,lx1*d); out=fma(ly0,r0,ly1*r1)
__device__ __forceinline__ float ac_blend(const BilinearTap& ty, const BilinearTap& tx, float a, float b, float c, float d) {
  float r0 = __fmaf_rn(tx.l0, a, __fmul_rn(tx.l1, b));
  float r1 = __fmaf_rn(tx.l0, c, __fmul_rn(tx.l1, d));
  return __fmaf_rn(ty.l0, r0, __fmul_rn(ty.l1, r1));
}
// ATen legacy nearest: min(int(floorf(dst * (float)in/out)), in-1)
__device__ __forceinline__ int nearest_src(int dst, float scale, int n_in) {
  return min((int)floorf(__fmul_rn((float)dst, scale)), n_in - 1);
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace prv2
