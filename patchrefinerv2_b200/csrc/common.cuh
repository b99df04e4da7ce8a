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

// (hi, lo) activation element access.  lo == nullptr -> plain bf16.
__device__ __forceinline__ float act_load(const bf16* hi, const bf16* lo, size_t i) {
  float v = bf2f(hi[i]);
  if (lo) v += bf2f(lo[i]);
  return v;
}
__device__ __forceinline__ void act_store(bf16* hi, bf16* lo, size_t i, float v) {
  bf16 h = f2bf(v);
  hi[i] = h;
  if (lo) lo[i] = f2bf(v - bf2f(h));
}

// 8-wide (16 byte) vector access
struct alignas(16) bf16x8 { bf16 v[8]; };

__device__ __forceinline__ void act_load8(const bf16* hi, const bf16* lo, size_t i, float (&out)[8]) {
  const uint4 a = *reinterpret_cast<const uint4*>(hi + i);
  const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(a2[k]); out[2 * k] = f.x; out[2 * k + 1] = f.y; }
  if (lo) {
    const uint4 b = *reinterpret_cast<const uint4*>(lo + i);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __bfloat1622float2(b2[k]); out[2 * k] += f.x; out[2 * k + 1] += f.y; }
  }
}
__device__ __forceinline__ void act_store8(bf16* hi, bf16* lo, size_t i, const float (&in)[8]) {
  // packed conversions (F2FP.BF16.PACK_AB: two values per instruction on the FMA pipe)
  __nv_bfloat162 h2[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) h2[k] = __floats2bfloat162_rn(in[2 * k], in[2 * k + 1]);
  *reinterpret_cast<uint4*>(hi + i) = *reinterpret_cast<const uint4*>(h2);
  if (lo) {
    __nv_bfloat162 l2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(h2[k]);
      l2[k] = __floats2bfloat162_rn(in[2 * k] - f.x, in[2 * k + 1] - f.y);
    }
    *reinterpret_cast<uint4*>(lo + i) = *reinterpret_cast<const uint4*>(l2);
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
