// Geometry kernels of the tiled-inference path: crop+resize, ROI gather, CAI blend.
// All HBM-bound gathers: one thread owns an output vector, reads are coalesced along the
// fastest output dimension, index math is replayed in fp32 exactly as the reference's
// CPU kernels do it (see oracle/pr_oracle.py np_* for the pinned restatements).
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <mutex>
#include <string.h>

namespace prv2 {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace prv2

using namespace prv2;

#ifndef PRV2_BUILD_DIGEST
#define PRV2_BUILD_DIGEST "unstamped"
#endif
extern "C" int prv2_version(void) { return PRV2_ABI_VERSION; }
static const char g_digest_marker[] = "PRV2_DIGEST=" PRV2_BUILD_DIGEST;       // build.py finds the stamp in the file by this marker
extern "C" const char* prv2_build_digest(void) { return g_digest_marker + 12; }
extern "C" const char* prv2_last_error(void) { return g_err; }
extern "C" int prv2_device_info(int32_t* out4) {
  PRV2_CHECK_ARG(out4 != nullptr, "prv2_device_info: null out");
  int dev = 0;
  PRV2_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  PRV2_CUDA(cudaGetDeviceProperties(&p, dev));
  out4[0] = p.multiProcessorCount;
  out4[1] = p.major;
  out4[2] = p.minor;
  out4[3] = (int32_t)p.sharedMemPerBlockOptin;
  return PRV2_OK;
}

// ---------------------------------------------------------------------------------------------
// The one real exchange of the sharded path: sum the packed partial buffer [num_canvas | m1_canvas | num_raw]
// over the ranks, in place, with NCCL on the caller's communicator and stream.  NCCL is resolved at run time from the process
// (PyTorch ships and loads libnccl.so.2), so the library has no link-time dependency on it and loads on a box without NCCL.
// ---------------------------------------------------------------------------------------------
#include <dlfcn.h>
namespace {
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int /*ncclDataType_t*/, int /*ncclRedOp_t*/, void* /*ncclComm_t*/, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
nccl_allreduce_fn g_nccl_allreduce = nullptr;
nccl_errstr_fn g_nccl_errstr = nullptr;
bool nccl_resolve() {
  if (g_nccl_allreduce) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // already mapped by the host framework?
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
  g_nccl_allreduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
  g_nccl_errstr = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
  return g_nccl_allreduce != nullptr;
}
}  // namespace

extern "C" int prv2_reduce_canvas(void* nccl_comm, float* packed, int64_t count, prv2_stream_t stream) {
  PRV2_CHECK_ARG(nccl_comm && packed && count >= 0, "prv2_reduce_canvas: null communicator / buffer");
  if (count == 0) return PRV2_OK;
  if (!nccl_resolve()) { set_error("prv2_reduce_canvas: libnccl.so.2 (ncclAllReduce) could not be resolved in this process"); return PRV2_ECUDA; }
  const int rc = g_nccl_allreduce(packed, packed, (size_t)count, 7 /*ncclFloat32*/, 0 /*ncclSum*/, nccl_comm, (cudaStream_t)stream);
  if (rc != 0) { set_error("prv2_reduce_canvas: ncclAllReduce failed (%d: %s)", rc, g_nccl_errstr ? g_nccl_errstr(rc) : "?"); return PRV2_ECUDA; }
  return PRV2_OK;
}

// ---------------------------------------------------------------------------------------------
// crop + bilinear(align_corners=True) resize            baseline_pretrain.py:272-280
// ---------------------------------------------------------------------------------------------
// grid: (ceil(pw/4/128), ph, P*3); each thread writes 4 consecutive x (one float4).
__global__ void __launch_bounds__(128) crop_resize_kernel(const float* __restrict__ image, int H, int W,
                                                          const int32_t* __restrict__ bboxs, float* __restrict__ out,
                                                          int ph, int pw) {
  const int p = blockIdx.z / 3, c = blockIdx.z % 3;
  const int y = blockIdx.y;
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (x4 >= pw) return;
  const int bx0 = bboxs[p * 4 + 0], by0 = bboxs[p * 4 + 1], bx1 = bboxs[p * 4 + 2], by1 = bboxs[p * 4 + 3];
  const int rw = bx1 - bx0, rh = by1 - by0;
  const float sy = ac_scale(rh, ph), sx = ac_scale(rw, pw);
  const BilinearTap ty = ac_tap(sy, y, rh);
  const float* r0 = image + ((size_t)c * H + (by0 + ty.i0)) * W + bx0;
  const float* r1 = image + ((size_t)c * H + (by0 + ty.i1)) * W + bx0;
  float v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x4 + k;
    if (x < pw) {
      const BilinearTap tx = ac_tap(sx, x, rw);
      v[k] = ac_blend(ty, tx, __ldg(r0 + tx.i0), __ldg(r0 + tx.i1), __ldg(r1 + tx.i0), __ldg(r1 + tx.i1));
    } else {
      v[k] = 0.f;
    }
  }
  float* o = out + (((size_t)p * 3 + c) * ph + y) * pw + x4;
  if (x4 + 3 < pw && (pw & 3) == 0) {
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    for (int k = 0; k < 4 && x4 + k < pw; ++k) o[k] = v[k];
  }
}

extern "C" int prv2_crop_resize(const float* image, int H, int W, const int32_t* bboxs, int P, float* out, int ph, int pw,
                                prv2_stream_t stream) {
  PRV2_CHECK_ARG(image && bboxs && out, "prv2_crop_resize: null pointer");
  PRV2_CHECK_ARG(H > 0 && W > 0 && ph > 0 && pw > 0 && P >= 0, "prv2_crop_resize: bad shape");
  if (P == 0) return PRV2_OK;
  PRV2_CHECK_ARG(P * 3 <= 65535 && ph <= 65535, "prv2_crop_resize: too many patches for one launch (%d)", P);
  dim3 grid(cdiv(cdiv(pw, 4), 128), ph, P * 3);
  crop_resize_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(image, H, W, bboxs, out, ph, pw);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

// ---------------------------------------------------------------------------------------------
// ROI gather (roi_align, aligned=True, one sample per bin)      patchrefiner.py:199-217
// ---------------------------------------------------------------------------------------------
struct RoiAxis { int lo, hi; float l, h; bool valid; };

// torchvision/csrc/ops/cpu/roi_align_common.h pre_calc_for_bilinear_interpolate, fp32, grid==1.
__device__ __forceinline__ RoiAxis roi_axis(float start, float bin, int p, int n_in) {
  RoiAxis a;
  float c = __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)), __fdiv_rn(__fmul_rn(0.5f, bin), 1.0f));
  a.valid = !(c < -1.0f || c > (float)n_in);
  c = fmaxf(c, 0.0f);
  int lo = (int)c;
  if (lo >= n_in - 1) {
    lo = n_in - 1;
    a.hi = lo;
    c = (float)lo;
  } else {
    a.hi = lo + 1;
  }
  a.lo = lo;
  a.l = __fsub_rn(c, (float)lo);
  a.h = __fsub_rn(1.0f, a.l);
  return a;
}

struct RoiGeom { float x1, y1, bw, bh; };
__device__ __forceinline__ RoiGeom roi_geom(const float* roi, float s, int h, int w) {
  RoiGeom g;
  g.x1 = __fsub_rn(__fmul_rn(roi[0], s), 0.5f);
  g.y1 = __fsub_rn(__fmul_rn(roi[1], s), 0.5f);
  float x2 = __fsub_rn(__fmul_rn(roi[2], s), 0.5f);
  float y2 = __fsub_rn(__fmul_rn(roi[3], s), 0.5f);
  g.bw = __fdiv_rn(__fsub_rn(x2, g.x1), (float)w);
  g.bh = __fdiv_rn(__fsub_rn(y2, g.y1), (float)h);
  return g;
}

__device__ __forceinline__ float roi_mix(const RoiAxis& ay, const RoiAxis& ax, float v1, float v2, float v3, float v4) {
  // w1*v1 + w2*v2 + w3*v3 + w4*v4, left to right, no contraction (torchvision CPU build)
  float w1 = __fmul_rn(ay.h, ax.h), w2 = __fmul_rn(ay.h, ax.l), w3 = __fmul_rn(ay.l, ax.h), w4 = __fmul_rn(ay.l, ax.l);
  float o = __fmul_rn(w1, v1);
  o = __fadd_rn(o, __fmul_rn(w2, v2));
  o = __fadd_rn(o, __fmul_rn(w3, v3));
  o = __fadd_rn(o, __fmul_rn(w4, v4));
  return o;
}

// one thread = one output pixel x 4 channels (float4), channels fastest -> coalesced both ways
__global__ void __launch_bounds__(256) roi_gather_f32_kernel(const float* __restrict__ feat, int h, int w, int C,
                                                             const float* __restrict__ rois, float s,
                                                             float* __restrict__ out, long long total) {
  const int cv = (C + 3) / 4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % cv) * 4;
    long long pix = idx / cv;
    const int x = (int)(pix % w);
    pix /= w;
    const int y = (int)(pix % h);
    const int p = (int)(pix / h);
    const RoiGeom g = roi_geom(rois + p * 4, s, h, w);
    const RoiAxis ay = roi_axis(g.y1, g.bh, y, h), ax = roi_axis(g.x1, g.bw, x, w);
    const bool valid = ay.valid && ax.valid;
    const float* f1 = feat + ((size_t)ay.lo * w + ax.lo) * C + c4;
    const float* f2 = feat + ((size_t)ay.lo * w + ax.hi) * C + c4;
    const float* f3 = feat + ((size_t)ay.hi * w + ax.lo) * C + c4;
    const float* f4 = feat + ((size_t)ay.hi * w + ax.hi) * C + c4;
    float* o = out + (((size_t)p * h + y) * w + x) * C + c4;
    for (int k = 0; k < 4 && c4 + k < C; ++k) o[k] = valid ? roi_mix(ay, ax, f1[k], f2[k], f3[k], f4[k]) : 0.f;
  }
}

// grid (ceil(w*C/8 / 256), ceil(h / ROI_ROWS), P).  A thread owns 8 channels (16 B of bf16) of one output column and walks
// ROI_ROWS consecutive output rows; the horizontal mix of a source row is computed once and reused by every output row
// that samples it (the ROI is up-sampled ~Sh x, so each source row serves several output rows).  The bf16 variant uses
// the separable form hy*(hx*v1 + lx*v2) + ly*(hx*v3 + lx*v4); only the f32 variant replays torchvision's rounding order.
constexpr int ROI_ROWS = 8;

template <bool X3>   // compile-time (hi, lo) planes: a runtime lo pointer leaves the second plane's code predicated off but issued
__global__ void __launch_bounds__(256) roi_gather_act_kernel(const bf16* __restrict__ fh, const bf16* __restrict__ fl_, int h, int w,
                                                             int C, int in_cs, const float* __restrict__ rois, float s,
                                                             bf16* __restrict__ oh, bf16* __restrict__ ol_, int out_cs) {
  const bf16* const fl = X3 ? fl_ : nullptr;
  bf16* const ol = X3 ? ol_ : nullptr;
  const unsigned cv = (unsigned)C >> 3, total = (unsigned)w * cv;
  const unsigned idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx / cv), c8 = (int)(idx % cv) * 8;
  const int y0 = blockIdx.y * ROI_ROWS, p = blockIdx.z;
  const RoiGeom g = roi_geom(rois + p * 4, s, h, w);
  const RoiAxis ax = roi_axis(g.x1, g.bw, x, w);
  auto hmix = [&](int row, float (&o)[8]) {
    float a[8], b[8];
    act_load8(fh, fl, ((size_t)row * w + ax.lo) * in_cs + c8, a);
    act_load8(fh, fl, ((size_t)row * w + ax.hi) * in_cs + c8, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaf(ax.h, a[k], ax.l * b[k]);
  };
  int r_lo = -1, r_hi = -1;
  float hb_lo[8], hb_hi[8];
#pragma unroll
  for (int yy = 0; yy < ROI_ROWS; ++yy) {
    const int y = y0 + yy;
    if (y >= h) break;
    const RoiAxis ay = roi_axis(g.y1, g.bh, y, h);
    if (ay.lo != r_lo) {
      if (ay.lo == r_hi) {
#pragma unroll
        for (int k = 0; k < 8; ++k) hb_lo[k] = hb_hi[k];
      } else {
        hmix(ay.lo, hb_lo);
      }
      r_lo = ay.lo;
    }
    if (ay.hi != r_hi) {
      if (ay.hi == r_lo) {
#pragma unroll
        for (int k = 0; k < 8; ++k) hb_hi[k] = hb_lo[k];
      } else {
        hmix(ay.hi, hb_hi);
      }
      r_hi = ay.hi;
    }
    const bool valid = ay.valid && ax.valid;
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = valid ? fmaf(ay.h, hb_lo[k], ay.l * hb_hi[k]) : 0.f;
    act_store8(oh, ol, (((size_t)p * h + y) * w + x) * out_cs + c8, o);
  }
}

static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;          // 16 resident 256-thread CTAs per SM, grid-stride beyond
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

extern "C" int prv2_roi_gather_f32(const float* feat, int h, int w, int C, const float* rois, int P, float spatial_scale, float* out,
                                   prv2_stream_t stream) {
  PRV2_CHECK_ARG(feat && rois && out, "prv2_roi_gather_f32: null pointer");
  PRV2_CHECK_ARG(h > 0 && w > 0 && C > 0 && P >= 0, "prv2_roi_gather_f32: bad shape");
  if (P == 0) return PRV2_OK;
  long long total = (long long)P * h * w * ((C + 3) / 4);
  roi_gather_f32_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(feat, h, w, C, rois, spatial_scale, out, total);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_roi_gather_act(const prv2_bf16* feat_hi, const prv2_bf16* feat_lo, int h, int w, int C, int in_cs, const float* rois,
                                   int P, float spatial_scale, prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs, prv2_stream_t stream) {
  PRV2_CHECK_ARG(feat_hi && rois && out_hi, "prv2_roi_gather_act: null pointer");
  PRV2_CHECK_ARG(h > 0 && w > 0 && C > 0 && P >= 0, "prv2_roi_gather_act: bad shape");
  PRV2_CHECK_ARG(C % 8 == 0 && in_cs % 8 == 0 && out_cs % 8 == 0, "prv2_roi_gather_act: C/pitch must be multiples of 8");
  if (P == 0) return PRV2_OK;
  PRV2_CHECK_ARG(h <= 65535 && P <= 65535, "prv2_roi_gather_act: grid too large");
  dim3 grid(cdiv((long long)w * (C / 8), 256), cdiv(h, ROI_ROWS), P);
  PRV2_CHECK_ARG((feat_lo == nullptr) == (out_lo == nullptr), "prv2_roi_gather_act: lo planes must both be present or absent");
  if (feat_lo)
    roi_gather_act_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)feat_hi, (const bf16*)feat_lo, h, w, C, in_cs, rois, spatial_scale,
                                                                       (bf16*)out_hi, (bf16*)out_lo, out_cs);
  else
    roi_gather_act_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)feat_hi, nullptr, h, w, C, in_cs, rois, spatial_scale,
                                                                        (bf16*)out_hi, nullptr, out_cs);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

// ---------------------------------------------------------------------------------------------
// CAI blend                                   estimator/models/utils.py:22-49
// ---------------------------------------------------------------------------------------------
#define PRV2_MAX_STAGES 8
struct StageTable { int n; prv2_grid_stage s[PRV2_MAX_STAGES]; };

// n / d for 0 <= n < 2^16 with the host-made magic m = floor(2^32 / d) + 1 (exact for d < 2^16): one IMAD.HI instead of the
// ~20-instruction generic division, per stage and thread, in a kernel that is otherwise four loads and a store
__device__ __forceinline__ int div_magic(int n, unsigned magic) { return (int)__umulhi((unsigned)n, magic); }
static unsigned make_magic(int d) { return (unsigned)((1ull << 32) / (unsigned)d) + 1u; }

// RunningAverageMap.update for one pixel, exactly as the tensor expression evaluates it:
// avg = (pred*ct + cnt*avg) / (cnt + ct); cnt = cnt + ct      (all separately rounded)
__device__ __forceinline__ void ram_update(float& avg, float& cnt, float pred, float ct) {
  float num = __fadd_rn(__fmul_rn(pred, ct), __fmul_rn(cnt, avg));
  float den = __fadd_rn(cnt, ct);
  avg = __fdiv_rn(num, den);
  cnt = den;
}

// Thread = 4 consecutive canvas pixels of one row: one (y / ph, x / pw) per stage per THREAD, float4
// loads of the mask / prediction rows when the group sits inside one patch, float4 stores.
// MODE 0: sequential-exact; 1: partial sums (own patches only); 2: finalize from reduced sums.
template <int MODE>
__global__ void __launch_bounds__(256) blend_canvas_kernel(const float* __restrict__ preds, const uint8_t* __restrict__ own,
                                                           const float* __restrict__ mask, int ph, int pw, StageTable st, int Hc,
                                                           int Wc, unsigned magic_h, unsigned magic_w, float* __restrict__ avg_out, float* __restrict__ cnt_out,
                                                           const float* __restrict__ num_in, const float* __restrict__ m1_in) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  if (x0 >= Wc) return;
  const size_t o = (size_t)y * Wc + x0;
  const bool vec_out = ((Wc & 3) == 0);
  float avg[4] = {0.f, 0.f, 0.f, 0.f}, cnt[4] = {0.f, 0.f, 0.f, 0.f}, num[4] = {0.f, 0.f, 0.f, 0.f}, m1[4] = {0.f, 0.f, 0.f, 0.f},
        cnt0[4] = {0.f, 0.f, 0.f, 0.f};
  if (MODE == 2) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (x0 + k < Wc) { num[k] = num_in[o + k]; m1[k] = m1_in[o + k]; }
  }
  for (int s = 0; s < st.n; ++s) {
    const prv2_grid_stage g = st.s[s];
    const int yy = y - g.off_h;
    if (yy < 0) continue;
    const int i = div_magic(yy, magic_h);
    if (i >= g.n_h) continue;
    const int ly = yy - i * ph;
    const int xx0 = x0 - g.off_w;
    float ct[4], pv[4];
    bool ok[4];
    int j0 = 0, lx0 = 0;
    if (xx0 >= 0) { j0 = div_magic(xx0, magic_w); lx0 = xx0 - j0 * pw; }
    const bool fast = xx0 >= 0 && j0 < g.n_w && lx0 + 3 < pw && x0 + 3 < Wc && ((lx0 | pw) & 3) == 0;
    if (fast) {
      const int pidx = g.first + i * g.n_w + j0;
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(mask + (size_t)ly * pw + lx0));
      ct[0] = c4.x; ct[1] = c4.y; ct[2] = c4.z; ct[3] = c4.w;
      const bool mine = (MODE == 0) || (MODE == 1 && own[pidx]);
      if (mine) {
        const float4 p4 = __ldg(reinterpret_cast<const float4*>(preds + ((size_t)pidx * ph + ly) * pw + lx0));
        pv[0] = p4.x; pv[1] = p4.y; pv[2] = p4.z; pv[3] = p4.w;
      } else {
        pv[0] = pv[1] = pv[2] = pv[3] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) ok[k] = (MODE != 1) || mine;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int xx = xx0 + k;
        ok[k] = false; ct[k] = 0.f; pv[k] = 0.f;
        if (xx < 0 || x0 + k >= Wc) continue;
        const int j = div_magic(xx, magic_w);
        if (j >= g.n_w) continue;
        const int lx = xx - j * pw, pidx = g.first + i * g.n_w + j;
        ct[k] = __ldg(mask + (size_t)ly * pw + lx);
        if (MODE == 2) { ok[k] = true; continue; }
        if (MODE == 1 && !own[pidx]) continue;
        pv[k] = __ldg(preds + ((size_t)pidx * ph + ly) * pw + lx);
        ok[k] = true;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!ok[k]) continue;
      if (MODE == 0) {
        if (s == 0) { avg[k] = pv[k]; cnt[k] = ct[k]; }                  // baseline_pretrain.py:352-355 (assignment)
        else if (ct[k] > 0.f) ram_update(avg[k], cnt[k], pv[k], ct[k]);   // utils.py:31-36
      } else if (MODE == 1) {
        if (s == 0) m1[k] = pv[k];
        else if (ct[k] > 0.f) num[k] = __fadd_rn(num[k], __fmul_rn(pv[k], ct[k]));
      } else {
        if (s == 0) { cnt[k] = ct[k]; cnt0[k] = ct[k]; }
        else if (ct[k] > 0.f) cnt[k] = __fadd_rn(cnt[k], ct[k]);
      }
    }
  }
  float r_a[4], r_c[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (MODE == 0) { r_a[k] = avg[k]; r_c[k] = cnt[k]; }
    else if (MODE == 1) { r_a[k] = num[k]; r_c[k] = m1[k]; }
    else {
      // sum form of the running mean: (m1*cnt0 + sum w p) / (cnt0 + sum w); untouched pixels keep m1
      r_a[k] = (cnt[k] > cnt0[k]) ? __fdiv_rn(__fadd_rn(__fmul_rn(m1[k], cnt0[k]), num[k]), cnt[k]) : m1[k];
      r_c[k] = cnt[k];
    }
  }
  if (MODE == 1) {       // accumulate into num_c (avg_out) and the m1 plane (cnt_out)
    for (int k = 0; k < 4 && x0 + k < Wc; ++k) { avg_out[o + k] += r_a[k]; cnt_out[o + k] += r_c[k]; }
  } else if (vec_out) {
    __stcs(reinterpret_cast<float4*>(avg_out + o), make_float4(r_a[0], r_a[1], r_a[2], r_a[3]));
    if (cnt_out) __stcs(reinterpret_cast<float4*>(cnt_out + o), make_float4(r_c[0], r_c[1], r_c[2], r_c[3]));
  } else {
    for (int k = 0; k < 4 && x0 + k < Wc; ++k) { avg_out[o + k] = r_a[k]; if (cnt_out) cnt_out[o + k] = r_c[k]; }
  }
}

// threads per CTA so that a canvas row splits into CTAs without a mostly idle last one (1792 / 4 = 448 = 4 x 112 -> 128 threads)
static int blend_threads(int Wc) {
  const int groups = cdiv(Wc, 4);
  static const char* cap_env = getenv("PRV2_BLEND_CANVAS_THREADS");        // A/B knob: threads per CTA cap
  const int cap = cap_env ? atoi(cap_env) : 128;                            // 128: 19.5 against 20.5 us by events at 256 (2160x3840, m2 canvas, cold L2); 14.9 us either way under ncu
  const int ctas = cdiv(groups, cap > 0 ? cap : 128);
  return cdiv(cdiv(groups, ctas), 32) * 32;
}

// Aligned fast path (pw, Wc and every stage's off_w multiples of 4 -- true for every tiling the reference produces):
// a 4-pixel group then lies inside exactly one patch of a stage or outside all of them, so a thread's whole working
// set is NS (mask float4, prediction float4) pairs.  All 2*NS loads are issued before the first dependent use
// (predicated, no branches between them): one DRAM latency per thread instead of one per stage.
template <int MODE, int NS>
__global__ void __launch_bounds__(256) blend_canvas_fast_kernel(const float* __restrict__ preds, const uint8_t* __restrict__ own,
                                                                const float* __restrict__ mask, int ph, int pw, StageTable st, int Hc,
                                                                int Wc, unsigned magic_h, unsigned magic_w, float* __restrict__ avg_out,
                                                                float* __restrict__ cnt_out, const float* __restrict__ num_in,
                                                                const float* __restrict__ m1_in) {
  const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  if (x0 >= Wc) return;
  const size_t o = (size_t)y * Wc + x0;
  float4 ct[NS], pv[NS];
  bool ok[NS];
  const int php = ph * pw;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const prv2_grid_stage g = st.s[s];
    // row part (the same for the whole CTA), kept apart from the column part; 32-bit element offsets (the host checks that
    // patches * ph * pw fits).  (Forming the row part once per CTA through shared memory saves another 8 % of the instructions and
    // no time: the barrier delays the loads of a kernel whose CTAs live for about one DRAM latency.)
    const int yy = y - g.off_h;
    const int i = div_magic(max(yy, 0), magic_h);
    const bool row_in = (s < st.n) && yy >= 0 && i < g.n_h;
    const int ly = yy - i * ph;
    const int row_first = g.first + i * g.n_w;                    // first patch of this row of patches
    const int m_row = ly * pw, p_row = (row_first * ph + ly) * pw;
    const int xx = x0 - g.off_w;
    const int j = div_magic(max(xx, 0), magic_w);
    const bool in = row_in && xx >= 0 && j < g.n_w;
    const int lx = xx - j * pw;
    bool mine = in;
    if (MODE == 1) mine = in && own[in ? row_first + j : 0] != 0;
    if (MODE == 2) mine = false;
    ct[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    pv[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in && (MODE != 1 || mine)) ct[s] = __ldg(reinterpret_cast<const float4*>(mask + (m_row + lx)));
    if (mine) pv[s] = __ldcs(reinterpret_cast<const float4*>(preds + (p_row + j * php + lx)));
    ok[s] = (MODE == 1) ? mine : in;
  }
  float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_c = make_float4(0.f, 0.f, 0.f, 0.f);
  if (MODE == 1) { acc_a = *reinterpret_cast<const float4*>(avg_out + o); acc_c = *reinterpret_cast<const float4*>(cnt_out + o); }
  if (MODE == 2) { acc_a = __ldcs(reinterpret_cast<const float4*>(num_in + o)); acc_c = __ldcs(reinterpret_cast<const float4*>(m1_in + o)); }
  float r_a[4], r_c[4];
  if (MODE == 0) {
    // stage-major: ONE branch per stage for the whole group (a group is inside one patch of a stage or outside all of them), the
    // weight test as a select -- utils.py:31-36 leaves a pixel untouched where the weight is not positive
    float avg[4], cnt[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { avg[k] = ok[0] ? (&pv[0].x)[k] : 0.f; cnt[k] = ok[0] ? (&ct[0].x)[k] : 0.f; }
#pragma unroll
    for (int s = 1; s < NS; ++s) {
      if (!ok[s]) continue;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float c = (&ct[s].x)[k], p = (&pv[s].x)[k];
        const float den = __fadd_rn(cnt[k], c);
        const float qv = __fdiv_rn(__fadd_rn(__fmul_rn(p, c), __fmul_rn(cnt[k], avg[k])), den);
        const bool pos = c > 0.f;
        avg[k] = pos ? qv : avg[k];
        cnt[k] = pos ? den : cnt[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { r_a[k] = avg[k]; r_c[k] = cnt[k]; }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float cnt = 0.f, num = (MODE == 2) ? (&acc_a.x)[k] : 0.f, m1 = (MODE == 2) ? (&acc_c.x)[k] : 0.f, cnt0 = 0.f;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const float c = (&ct[s].x)[k], p = (&pv[s].x)[k];
        if (!ok[s]) continue;
        if (MODE == 1) {
          if (s == 0) m1 = p;
          else if (c > 0.f) num = __fadd_rn(num, __fmul_rn(p, c));
        } else {
          if (s == 0) { cnt = c; cnt0 = c; }
          else if (c > 0.f) cnt = __fadd_rn(cnt, c);
        }
      }
      if (MODE == 1) { r_a[k] = (&acc_a.x)[k] + num; r_c[k] = (&acc_c.x)[k] + m1; }
      else { r_a[k] = (cnt > cnt0) ? __fdiv_rn(__fadd_rn(__fmul_rn(m1, cnt0), num), cnt) : m1; r_c[k] = cnt; }
    }
  }
  if (MODE == 1) {
    *reinterpret_cast<float4*>(avg_out + o) = make_float4(r_a[0], r_a[1], r_a[2], r_a[3]);
    *reinterpret_cast<float4*>(cnt_out + o) = make_float4(r_c[0], r_c[1], r_c[2], r_c[3]);
  } else {
    __stcs(reinterpret_cast<float4*>(avg_out + o), make_float4(r_a[0], r_a[1], r_a[2], r_a[3]));
    if (cnt_out) __stcs(reinterpret_cast<float4*>(cnt_out + o), make_float4(r_c[0], r_c[1], r_c[2], r_c[3]));
  }
}

// test / A-B hook: route every blend entry point through the generic (any-alignment) kernels
static int g_blend_generic = -1;
static bool blend_generic_forced() {
  if (g_blend_generic < 0) { const char* e = getenv("PRV2_BLEND_GENERIC"); g_blend_generic = (e && e[0] == '1') ? 1 : 0; }
  return g_blend_generic == 1;
}
static void seg_knobs(int on);
static int g_seg_launches = 0;                                   // (test hook: how often the segment kernel took a *_raw call)
extern "C" int prv2_debug_blend_generic(int on) {
  if (on == 0x10000) return g_seg_launches;                      // query only
  g_blend_generic = (on & 1) ? 1 : 0;
  seg_knobs(on);
  return PRV2_OK;
}

static bool canvas_aligned(const StageTable& t, int ph, int pw, int Wc) {
  if (blend_generic_forced() || (pw & 3) || (Wc & 3) || t.n > 4) return false;
  for (int i = 0; i < t.n; ++i) {
    if (t.s[i].off_w & 3) return false;
    if (((long long)t.s[i].first + (long long)t.s[i].n_h * t.s[i].n_w + 1) * ph * pw >= (1ll << 31)) return false;      // 32-bit element offsets in the fast kernel
  }
  return true;
}

template <int MODE>
static void launch_canvas(const float* preds, const uint8_t* own, const float* mask, int ph, int pw, const StageTable& t, int Hc, int Wc,
                          float* a, float* b, const float* num_in, const float* m1_in, cudaStream_t stream) {
  const int bt = blend_threads(Wc);
  dim3 grid(cdiv(cdiv(Wc, 4), bt), Hc);
  const unsigned mh = make_magic(ph), mw = make_magic(pw);
  if (canvas_aligned(t, ph, pw, Wc)) {
    if (t.n == 1) blend_canvas_fast_kernel<MODE, 1><<<grid, bt, 0, stream>>>(preds, own, mask, ph, pw, t, Hc, Wc, mh, mw, a, b, num_in, m1_in);
    else if (t.n == 2) blend_canvas_fast_kernel<MODE, 2><<<grid, bt, 0, stream>>>(preds, own, mask, ph, pw, t, Hc, Wc, mh, mw, a, b, num_in, m1_in);
    else blend_canvas_fast_kernel<MODE, 4><<<grid, bt, 0, stream>>>(preds, own, mask, ph, pw, t, Hc, Wc, mh, mw, a, b, num_in, m1_in);
  } else {
    blend_canvas_kernel<MODE><<<grid, bt, 0, stream>>>(preds, own, mask, ph, pw, t, Hc, Wc, mh, mw, a, b, num_in, m1_in);
  }
}

static int fill_stages(StageTable& t, const prv2_grid_stage* stages, int n) {
  if (!stages || n < 1 || n > PRV2_MAX_STAGES) return -1;
  t.n = n;
  for (int i = 0; i < n; ++i) t.s[i] = stages[i];
  return 0;
}

extern "C" int prv2_blend_canvas(const float* preds, const float* mask, int ph, int pw, const prv2_grid_stage* stages, int n_stages,
                                 int Hc, int Wc, float* avg, float* cnt, prv2_stream_t stream) {
  PRV2_CHECK_ARG(preds && mask && avg, "prv2_blend_canvas: null pointer");
  StageTable t;
  PRV2_CHECK_ARG(fill_stages(t, stages, n_stages) == 0, "prv2_blend_canvas: need 1..%d stages", PRV2_MAX_STAGES);
  PRV2_CHECK_ARG(ph > 0 && pw > 0 && Hc > 0 && Wc > 0 && Hc <= 65535 && Wc <= 65535 && ph < 65536 && pw < 65536, "prv2_blend_canvas: bad shape");
  launch_canvas<0>(preds, nullptr, mask, ph, pw, t, Hc, Wc, avg, cnt, nullptr, nullptr, (cudaStream_t)stream);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_blend_partial_canvas(const float* preds, const uint8_t* own, const float* mask, int ph, int pw,
                                         const prv2_grid_stage* stages, int n_stages, int Hc, int Wc, float* num_c, float* m1,
                                         prv2_stream_t stream) {
  PRV2_CHECK_ARG(preds && own && mask && num_c && m1, "prv2_blend_partial_canvas: null pointer");
  StageTable t;
  PRV2_CHECK_ARG(fill_stages(t, stages, n_stages) == 0, "prv2_blend_partial_canvas: need 1..%d stages", PRV2_MAX_STAGES);
  PRV2_CHECK_ARG(ph > 0 && pw > 0 && Hc > 0 && Wc > 0 && Hc <= 65535, "prv2_blend_partial_canvas: bad shape");
  launch_canvas<1>(preds, own, mask, ph, pw, t, Hc, Wc, num_c, m1, nullptr, nullptr, (cudaStream_t)stream);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_blend_finalize_canvas(const float* num_c, const float* m1, const float* mask, int ph, int pw,
                                          const prv2_grid_stage* stages, int n_stages, int Hc, int Wc, float* avg, float* cnt,
                                          prv2_stream_t stream) {
  PRV2_CHECK_ARG(num_c && m1 && mask && avg, "prv2_blend_finalize_canvas: null pointer");
  StageTable t;
  PRV2_CHECK_ARG(fill_stages(t, stages, n_stages) == 0, "prv2_blend_finalize_canvas: need 1..%d stages", PRV2_MAX_STAGES);
  PRV2_CHECK_ARG(ph > 0 && pw > 0 && Hc > 0 && Wc > 0 && Hc <= 65535, "prv2_blend_finalize_canvas: bad shape");
  launch_canvas<2>(nullptr, nullptr, mask, ph, pw, t, Hc, Wc, avg, cnt, num_c, m1, (cudaStream_t)stream);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

// rN stage.  Block = one raw row segment of 1024 pixels, thread = 4 consecutive pixels.  Warp 0
// compacts the random-patch list to the patches covering this row (ballot, draw order kept) and
// precomputes their row offsets; every pixel then walks that short list in order.  All resampling
// scales are fp32 quotients formed once on the host (same IEEE division the device would do).
#define PRV2_MAX_RANDOM 512
struct RawScales { float ns_y, ns_x, bs_y, bs_x, ps_y, ps_x; };

template <int MODE>
__global__ void __launch_bounds__(256) blend_raw_kernel(const float* __restrict__ avg_c, const float* __restrict__ cnt_c, int Hc, int Wc,
                                                        const float* __restrict__ preds, const uint8_t* __restrict__ own,
                                                        const int32_t* __restrict__ starts, int n, int ph, int pw,
                                                        const float* __restrict__ rmask, int rh, int rw, int H, int W,
                                                        float* __restrict__ out, float* __restrict__ out_cnt,
                                                        const float* __restrict__ num_in, RawScales sc) {
  __shared__ int s_x0[PRV2_MAX_RANDOM], s_prow[PRV2_MAX_RANDOM], s_mrow[PRV2_MAX_RANDOM];
  __shared__ int s_n;
  const int y = blockIdx.y;
  if (threadIdx.x < 32) {
    int m = 0;
    for (int base = 0; base < n; base += 32) {
      const int k = base + threadIdx.x;
      int y0 = 0, x0 = 0;
      bool hit = false;
      if (k < n) {
        y0 = starts[k * 2 + 0]; x0 = starts[k * 2 + 1];
        hit = (y >= y0 && y < y0 + rh) && (MODE != 1 || own[k]);
      }
      const unsigned b = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int pos = m + __popc(b & ((1u << threadIdx.x) - 1));
        const int ly = y - y0;
        s_x0[pos] = x0;
        s_mrow[pos] = ly * rw;
        // baseline_pretrain.py:210 F.interpolate(predictions, patch_raw_shape) (nearest), row part
        s_prow[pos] = (k * ph + nearest_src(ly, sc.ps_y, ph)) * pw;
      }
      m += __popc(b);
    }
    if (threadIdx.x == 0) s_n = m;
  }
  __syncthreads();
  const int m = s_n;
  // one CTA walks its whole row (the covering-patch list above is built once per row, not once per 1024 pixels)
  for (int xb = threadIdx.x * 4; xb < W; xb += blockDim.x * 4) {
  const size_t o = (size_t)y * W + xb;
  float avg[4] = {0.f, 0.f, 0.f, 0.f}, cnt[4] = {0.f, 0.f, 0.f, 0.f}, num[4] = {0.f, 0.f, 0.f, 0.f}, c0[4] = {0.f, 0.f, 0.f, 0.f};
  if (MODE != 1) {
    // utils.py:42: average map -> nearest ; utils.py:43: count map -> bilinear(align_corners=True)
    const float* arow = avg_c + (size_t)nearest_src(y, sc.ns_y, Hc) * Wc;
    const BilinearTap ty = ac_tap(sc.bs_y, y, Hc);
    const float* c_r0 = cnt_c + (size_t)ty.i0 * Wc;
    const float* c_r1 = cnt_c + (size_t)ty.i1 * Wc;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int x = xb + q;
      if (x >= W) continue;
      avg[q] = __ldg(arow + nearest_src(x, sc.ns_x, Wc));
      const BilinearTap tx = ac_tap(sc.bs_x, x, Wc);
      cnt[q] = ac_blend(ty, tx, __ldg(c_r0 + tx.i0), __ldg(c_r0 + tx.i1), __ldg(c_r1 + tx.i0), __ldg(c_r1 + tx.i1));
      c0[q] = cnt[q];
    }
    if (MODE == 2) {
#pragma unroll
      for (int q = 0; q < 4; ++q) if (xb + q < W) num[q] = num_in[o + q];
    }
  }
  for (int i = 0; i < m; ++i) {
    const int x0 = s_x0[i];
    if (xb + 3 < x0 || xb >= x0 + rw) continue;            // group entirely outside this patch
    const float* mrow = rmask + s_mrow[i];
    const float* prow = preds + s_prow[i];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int lx = xb + q - x0;
      if (lx < 0 || lx >= rw || xb + q >= W) continue;
      const float ct = __ldg(mrow + lx);
      if (!(ct > 0.f)) continue;
      if (MODE == 2) { cnt[q] = __fadd_rn(cnt[q], ct); continue; }
      const float p = __ldg(prow + nearest_src(lx, sc.ps_x, pw));
      if (MODE == 0) ram_update(avg[q], cnt[q], p, ct);
      else num[q] = __fadd_rn(num[q], __fmul_rn(p, ct));
    }
  }
  float r_a[4], r_c[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (MODE == 0) { r_a[q] = avg[q]; r_c[q] = cnt[q]; }
    else if (MODE == 1) { r_a[q] = num[q]; r_c[q] = 0.f; }
    else { r_a[q] = (cnt[q] > c0[q]) ? __fdiv_rn(__fadd_rn(__fmul_rn(avg[q], c0[q]), num[q]), cnt[q]) : avg[q]; r_c[q] = cnt[q]; }
  }
  if (MODE == 1) {
    for (int q = 0; q < 4 && xb + q < W; ++q) out[o + q] += r_a[q];
  } else if ((W & 3) == 0) {
    __stcs(reinterpret_cast<float4*>(out + o), make_float4(r_a[0], r_a[1], r_a[2], r_a[3]));
    if (out_cnt) __stcs(reinterpret_cast<float4*>(out_cnt + o), make_float4(r_c[0], r_c[1], r_c[2], r_c[3]));
  } else {
    for (int q = 0; q < 4 && xb + q < W; ++q) { out[o + q] = r_a[q]; if (out_cnt) out_cnt[o + q] = r_c[q]; }
  }
  }
}

static RawScales raw_scales(int Hc, int Wc, int H, int W, int ph, int pw, int rh, int rw) {
  RawScales s;
  s.ns_y = (float)Hc / (float)H; s.ns_x = (float)Wc / (float)W;                  // ATen nearest: (float)in / out
  s.bs_y = H > 1 ? (float)(Hc - 1) / (float)(H - 1) : 0.f;                       // ATen bilinear align_corners
  s.bs_x = W > 1 ? (float)(Wc - 1) / (float)(W - 1) : 0.f;
  s.ps_y = (float)ph / (float)rh; s.ps_x = (float)pw / (float)rw;
  return s;
}

// Table-driven version of the rN stage (same CTA shape as blend_raw_kernel: one raw row per CTA, thread = 4 consecutive
// pixels).  Everything that depends on the column only -- nearest column of the average canvas, bilinear tap (i0, l1) of
// the count canvas, nearest column of a prediction for every in-patch x -- is read from small device tables that
// prv2_blend_raw_* builds once per geometry (raw_tables_kernel), instead of being recomputed with int<->float conversions for
// each of the H rows.  The arithmetic on pixel values is unchanged (same ops, same order) -> same bits as blend_raw_kernel.
struct RawTables { const unsigned short* a_ix; const unsigned short* c_i0; const float* c_l1; const unsigned short* src_x; const int* a_b; const int* c_b; };
// Prepared patch mask (prv2_blend_raw_prepare): four copies of the [rh, rw] weight map, copy c shifted right by c columns and
// zero-padded to `pitch` (a multiple of 4, >= rw + 8), and four equally shifted copies of the nearest-source-column table.  For
// an output group starting l0 columns into a patch, copy c = (-l0) & 3 holds the group's four weights / source columns in ONE
// aligned 16-byte / 8-byte vector -- also where the group straddles a patch edge (the padding weighs 0).
struct RawPrep { const float* mask4; const unsigned short* srcx4; const int* srci4; int pitch; };

__global__ void raw_tables_kernel(unsigned short* a_ix, unsigned short* c_i0, float* c_l1, unsigned short* src_x, int* a_b, int* c_b, int Wc, int W,
                                  int Wpad, int pw, int rw, RawScales sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Wpad) {
    const int x = min(i, W - 1);
    const int a = nearest_src(x, sc.ns_x, Wc);
    a_ix[i] = (unsigned short)a;
    const BilinearTap tx = ac_tap(sc.bs_x, x, Wc);
    c_i0[i] = (unsigned short)tx.i0;
    c_l1[i] = tx.l1;
    a_b[i] = a * 4; c_b[i] = tx.i0 * 4;                           // the same columns as byte offsets (segment kernel)
  }
  if (i < rw) src_x[i] = (unsigned short)nearest_src(i, sc.ps_x, pw);
}

template <int MODE>
__global__ void __launch_bounds__(256) blend_raw_tab_kernel(const float* __restrict__ avg_c, const float* __restrict__ cnt_c, int Hc, int Wc,
                                                            const float* __restrict__ preds, const uint8_t* __restrict__ own,
                                                            const int32_t* __restrict__ starts, int n, int ph, int pw,
                                                            const float* __restrict__ rmask, int rh, int rw, int H, int W,
                                                            float* __restrict__ out, float* __restrict__ out_cnt,
                                                            const float* __restrict__ num_in, RawScales sc, RawTables tb, RawPrep prep) {
  __shared__ int s_x0[PRV2_MAX_RANDOM], s_prow[PRV2_MAX_RANDOM], s_mrow[PRV2_MAX_RANDOM];
  __shared__ int s_n;
  const int y = blockIdx.y;
  if (threadIdx.x < 32) {
    int m = 0;
    for (int base = 0; base < n; base += 32) {
      const int k = base + threadIdx.x;
      int y0 = 0, x0 = 0;
      bool hit = false;
      if (k < n) {
        y0 = starts[k * 2 + 0]; x0 = starts[k * 2 + 1];
        hit = (y >= y0 && y < y0 + rh) && (MODE != 1 || own[k]);
      }
      const unsigned b = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int pos = m + __popc(b & ((1u << threadIdx.x) - 1));
        const int ly = y - y0;
        s_x0[pos] = x0;
        s_mrow[pos] = ly * (prep.mask4 ? prep.pitch : rw);
        s_prow[pos] = (MODE == 2) ? 0 : (k * ph + nearest_src(ly, sc.ps_y, ph)) * pw;    // baseline_pretrain.py:210 (nearest), row part
      }
      m += __popc(b);
    }
    if (threadIdx.x == 0) s_n = m;
  }
  // row-only resampling state (utils.py:42-43), uniform over the CTA
  const float* arow = nullptr; const float* c_r0 = nullptr; const float* c_r1 = nullptr;
  BilinearTap ty; ty.i0 = ty.i1 = 0; ty.l0 = ty.l1 = 0.f;
  if (MODE != 1) {
    arow = avg_c + (size_t)nearest_src(y, sc.ns_y, Hc) * Wc;
    ty = ac_tap(sc.bs_y, y, Hc);
    c_r0 = cnt_c + (size_t)ty.i0 * Wc;
    c_r1 = cnt_c + (size_t)ty.i1 * Wc;
  }
  __syncthreads();
  const int m = s_n;
  const bool vec = (W & 3) == 0;
  for (int xb = threadIdx.x * 4; xb < W; xb += blockDim.x * 4) {
    const size_t o = (size_t)y * W + xb;
    const bool full = xb + 3 < W;
    float avg[4] = {0.f, 0.f, 0.f, 0.f}, cnt[4] = {0.f, 0.f, 0.f, 0.f}, num[4] = {0.f, 0.f, 0.f, 0.f}, c0[4] = {0.f, 0.f, 0.f, 0.f};
    if (MODE != 1) {
      // tables are padded to a multiple of 4 columns (clamped to W-1), so the vector loads are always in bounds
      const uint2 ta = __ldg(reinterpret_cast<const uint2*>(tb.a_ix + xb));
      const uint2 ti = __ldg(reinterpret_cast<const uint2*>(tb.c_i0 + xb));
      const float4 tl = __ldg(reinterpret_cast<const float4*>(tb.c_l1 + xb));
      const int a_ix[4] = {(int)(ta.x & 0xffffu), (int)(ta.x >> 16), (int)(ta.y & 0xffffu), (int)(ta.y >> 16)};
      const int i0[4] = {(int)(ti.x & 0xffffu), (int)(ti.x >> 16), (int)(ti.y & 0xffffu), (int)(ti.y >> 16)};
      float va[4], v00[4], v01[4], v10[4], v11[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) va[q] = __ldg(arow + a_ix[q]);
      if (i0[3] - i0[0] <= 2) {
        // up-sampling: the four outputs' taps (i0, i0 + 1) lie in FOUR consecutive canvas columns -- 8 loads instead of 16
        float r0[4], r1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int col = min(i0[0] + k, Wc - 1);               // (i1 clamps the same way at the right edge)
          r0[k] = __ldg(c_r0 + col); r1[k] = __ldg(c_r1 + col);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int d = i0[q] - i0[0];
          v00[q] = d == 0 ? r0[0] : d == 1 ? r0[1] : r0[2];
          v01[q] = d == 0 ? r0[1] : d == 1 ? r0[2] : r0[3];
          v10[q] = d == 0 ? r1[0] : d == 1 ? r1[1] : r1[2];
          v11[q] = d == 0 ? r1[1] : d == 1 ? r1[2] : r1[3];
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i1 = i0[q] + (i0[q] < Wc - 1 ? 1 : 0);
          v00[q] = __ldg(c_r0 + i0[q]); v01[q] = __ldg(c_r0 + i1);
          v10[q] = __ldg(c_r1 + i0[q]); v11[q] = __ldg(c_r1 + i1);
        }
      }
      float4 nin = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 2) {
        if (full && vec) nin = __ldcs(reinterpret_cast<const float4*>(num_in + o));
        else { for (int q = 0; q < 4; ++q) if (xb + q < W) (&nin.x)[q] = num_in[o + q]; }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        BilinearTap tx;
        tx.l1 = (&tl.x)[q]; tx.l0 = __fsub_rn(1.0f, tx.l1);
        avg[q] = va[q];
        cnt[q] = ac_blend(ty, tx, v00[q], v01[q], v10[q], v11[q]);
        c0[q] = cnt[q];
        num[q] = (&nin.x)[q];
      }
    }
    for (int i = 0; i < m; ++i) {
      const int l0 = xb - s_x0[i];
      if (l0 + 3 < 0 || l0 >= rw) continue;                    // group entirely outside this patch
      const float* mrow = prep.mask4 ? prep.mask4 + s_mrow[i] : rmask + s_mrow[i];      // copy 0 of the prepared mask is unshifted
      const float* prow = preds + s_prow[i];
      if (prep.mask4 && full) {
        // one aligned float4 of weights and one uint2 of source columns from the copy shifted by (-l0) & 3 (edge groups too)
        const int c = (-l0) & 3, j = l0 + c;
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(prep.mask4 + (size_t)c * rh * prep.pitch + s_mrow[i] + j));
        const float ct[4] = {c4.x, c4.y, c4.z, c4.w};
        float p[4] = {0.f, 0.f, 0.f, 0.f};
        if (MODE != 2) {
          const uint2 s2 = __ldg(reinterpret_cast<const uint2*>(prep.srcx4 + (size_t)c * prep.pitch + j));
          const int sx[4] = {(int)(s2.x & 0xffffu), (int)(s2.x >> 16), (int)(s2.y & 0xffffu), (int)(s2.y >> 16)};
#pragma unroll
          for (int q = 0; q < 4; ++q) p[q] = __ldg(prow + sx[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (!(ct[q] > 0.f)) continue;
          if (MODE == 2) cnt[q] = __fadd_rn(cnt[q], ct[q]);
          else if (MODE == 0) ram_update(avg[q], cnt[q], p[q], ct[q]);
          else num[q] = __fadd_rn(num[q], __fmul_rn(p[q], ct[q]));
        }
      } else if (!prep.mask4 && l0 >= 0 && l0 + 3 < rw && full) {
        float ct[4], p[4];
        int sx[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { ct[q] = __ldg(mrow + l0 + q); sx[q] = (MODE != 2) ? (int)__ldg(tb.src_x + l0 + q) : 0; }
        if (MODE != 2) {
#pragma unroll
          for (int q = 0; q < 4; ++q) p[q] = __ldg(prow + sx[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (!(ct[q] > 0.f)) continue;
          if (MODE == 2) cnt[q] = __fadd_rn(cnt[q], ct[q]);
          else if (MODE == 0) ram_update(avg[q], cnt[q], p[q], ct[q]);
          else num[q] = __fadd_rn(num[q], __fmul_rn(p[q], ct[q]));
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int lx = l0 + q;
          if (lx < 0 || lx >= rw || xb + q >= W) continue;
          const float ct = __ldg(mrow + lx);
          if (!(ct > 0.f)) continue;
          if (MODE == 2) { cnt[q] = __fadd_rn(cnt[q], ct); continue; }
          const float p = __ldg(prow + (int)__ldg(tb.src_x + lx));
          if (MODE == 0) ram_update(avg[q], cnt[q], p, ct);
          else num[q] = __fadd_rn(num[q], __fmul_rn(p, ct));
        }
      }
    }
    float r_a[4], r_c[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (MODE == 0) { r_a[q] = avg[q]; r_c[q] = cnt[q]; }
      else if (MODE == 1) { r_a[q] = num[q]; r_c[q] = 0.f; }
      else { r_a[q] = (cnt[q] > c0[q]) ? __fdiv_rn(__fadd_rn(__fmul_rn(avg[q], c0[q]), num[q]), cnt[q]) : avg[q]; r_c[q] = cnt[q]; }
    }
    if (MODE == 1) {
      for (int q = 0; q < 4 && xb + q < W; ++q) out[o + q] += r_a[q];
    } else if (full && vec) {
      __stcs(reinterpret_cast<float4*>(out + o), make_float4(r_a[0], r_a[1], r_a[2], r_a[3]));
      if (out_cnt) __stcs(reinterpret_cast<float4*>(out_cnt + o), make_float4(r_c[0], r_c[1], r_c[2], r_c[3]));
    } else {
      for (int q = 0; q < 4 && xb + q < W; ++q) { out[o + q] = r_a[q]; if (out_cnt) out_cnt[o + q] = r_c[q]; }
    }
  }
}

// Column tables live in a small per-geometry device cache owned by the library (<= 64 KB each, built by one tiny kernel the
// first time a geometry is seen on a device; an event orders later use from other streams after the build).
struct RawTableEntry {
  int dev, Wc, W, pw, rw, used;
  void* mem;
  cudaEvent_t ready;
  RawTables tb;
};
#define PRV2_RAW_TABLE_SLOTS 64
static RawTableEntry g_raw_tables[PRV2_RAW_TABLE_SLOTS];
static std::mutex g_raw_tables_mu;                               // entry points may be called from several host threads

static bool raw_tables_get(int Wc, int W, int pw, int rw, const RawScales& sc, cudaStream_t stream, RawTables* tb) {
  if (Wc > 65535 || pw > 65535 || W <= 0) return false;
  std::lock_guard<std::mutex> lock(g_raw_tables_mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  RawTableEntry* slot = nullptr;
  for (int i = 0; i < PRV2_RAW_TABLE_SLOTS; ++i) {
    RawTableEntry& e = g_raw_tables[i];
    if (e.used && e.dev == dev && e.Wc == Wc && e.W == W && e.pw == pw && e.rw == rw) {
      if (cudaStreamWaitEvent(stream, e.ready, 0) != cudaSuccess) return false;
      *tb = e.tb;
      return true;
    }
    if (!e.used && !slot) slot = &e;
  }
  if (!slot) return false;                                      // cache full: the caller falls back to the generic kernel
  const int Wpad = (W + 3) & ~3, rwpad = (rw + 7) & ~7;
  const size_t bytes = (size_t)Wpad * (2 + 2 + 4 + 4 + 4) + (size_t)rwpad * 2 + 64;
  void* mem = nullptr;
  if (cudaMalloc(&mem, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
  char* base = (char*)mem;
  float* c_l1 = (float*)base;                                   // 16-byte aligned first
  int* a_b = (int*)(base + (size_t)Wpad * 4);
  int* c_b = a_b + Wpad;
  unsigned short* a_ix = (unsigned short*)(c_b + Wpad);
  unsigned short* c_i0 = a_ix + Wpad;
  unsigned short* src_x = c_i0 + Wpad;
  const int nthr = Wpad > rw ? Wpad : rw;
  raw_tables_kernel<<<cdiv(nthr, 256), 256, 0, stream>>>(a_ix, c_i0, c_l1, src_x, a_b, c_b, Wc, W, Wpad, pw, rw, sc);
  cudaEvent_t ev;
  if (cudaGetLastError() != cudaSuccess || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { cudaFree(mem); return false; }
  cudaEventRecord(ev, stream);
  slot->dev = dev; slot->Wc = Wc; slot->W = W; slot->pw = pw; slot->rw = rw; slot->mem = mem; slot->ready = ev;
  slot->tb.a_ix = a_ix; slot->tb.c_i0 = c_i0; slot->tb.c_l1 = c_l1; slot->tb.src_x = src_x; slot->tb.a_b = a_b; slot->tb.c_b = c_b;
  slot->used = 1;
  *tb = slot->tb;
  return true;
}

static int raw_prep_pitch(int rw) { return ((rw + 3) & ~3) + 8; }
static size_t raw_prep_bytes(int rh, int rw) {                   // four shifted weight maps | source columns (u16) | the same as i32
  return (size_t)4 * rh * raw_prep_pitch(rw) * sizeof(float) + (size_t)4 * raw_prep_pitch(rw) * (sizeof(unsigned short) + sizeof(int));
}

__global__ void raw_prepare_kernel(const float* __restrict__ rmask, int rh, int rw, int pw, int pitch, float ps_x, float* __restrict__ mask4,
                                   unsigned short* __restrict__ srcx4, int* __restrict__ srci4) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y, c = blockIdx.z;
  if (j >= pitch) return;
  const int src = j - c;                                         // copy c is shifted right by c columns, zero elsewhere
  const bool in = src >= 0 && src < rw;
  mask4[((size_t)c * rh + row) * pitch + j] = in ? rmask[(size_t)row * rw + src] : 0.f;
  if (row == 0) {
    // the padding weighs 0; its source column is the nearest real one, so that a padded pixel still reads inside the run
    const int sx = nearest_src(min(max(src, 0), rw - 1), ps_x, pw);
    srcx4[(size_t)c * pitch + j] = (unsigned short)sx;
    srci4[(size_t)c * pitch + j] = sx;
  }
}

static RawPrep raw_prep_view(const void* prep, int rh, int rw) {
  RawPrep v;
  v.mask4 = nullptr; v.srcx4 = nullptr; v.srci4 = nullptr; v.pitch = 0;
  if (prep) {
    v.pitch = raw_prep_pitch(rw);
    v.mask4 = (const float*)prep;
    v.srcx4 = (const unsigned short*)((const char*)prep + (size_t)4 * rh * v.pitch * sizeof(float));
    v.srci4 = (const int*)(v.srcx4 + (size_t)4 * v.pitch);                            // 8 * pitch bytes further: 16-byte aligned (pitch % 4 == 0)
  }
  return v;
}

extern "C" int64_t prv2_blend_raw_prep_bytes(int rh, int rw) { return (rh > 0 && rw > 0) ? (int64_t)raw_prep_bytes(rh, rw) : 0; }

extern "C" int prv2_blend_raw_prepare(const float* rmask, int rh, int rw, int pw, void* prep, prv2_stream_t stream) {
  PRV2_CHECK_ARG(rmask && prep, "prv2_blend_raw_prepare: null pointer");
  PRV2_CHECK_ARG(rh > 0 && rw > 0 && pw > 0 && pw <= 65535 && rh <= 65535, "prv2_blend_raw_prepare: bad shape");
  PRV2_CHECK_ARG(((uintptr_t)prep & 15) == 0, "prv2_blend_raw_prepare: the buffer must be 16-byte aligned");
  const int pitch = raw_prep_pitch(rw);
  const RawPrep v = raw_prep_view(prep, rh, rw);
  raw_prepare_kernel<<<dim3(cdiv(pitch, 256), rh, 4), 256, 0, (cudaStream_t)stream>>>(rmask, rh, rw, pw, pitch, (float)pw / (float)rw, (float*)v.mask4,
                                                                                      (unsigned short*)v.srcx4, (int*)v.srci4);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}


// ---------------------------------------------------------------------------------------------
// Segment kernel of the rN stage (MODE 0: sequential-exact, MODE 2: finalize from reduced sums).  Same arithmetic on pixel values,
// in the same order, as blend_raw_kernel / blend_raw_tab_kernel -> same bits; what changes is how the operands arrive:
//  * CTA = one raw row x one x segment of 128*w pixels (w warps, thread = 4 consecutive pixels): many small CTAs, ~2000 threads
//    per SM, so that the two dependent L2 reads per covering patch (weights, then the up-sampled prediction) are hidden.
//  * The canvas rows the raw row needs (nearest row of the average map, the two bilinear rows of the count map), restricted to
//    the segment's column span, are staged in shared memory by three bulk copies (cp.async.bulk + mbarrier) issued by one thread:
//    the 20 canvas reads per group become LDS with immediate offsets from per-column byte offsets read as vectors from tables.
//  * Warp 0 lists the patches covering (row, segment) in draw order as ready-made operands: with c = x0 & 3 the group at absolute
//    column xb finds its four weights at mask4 + moff + xb (copy c of the prepared weight map) and its four source columns at
//    srci4 + soff + xb, for EVERY group of the row (c does not depend on the group because xb is a multiple of 4).
// (A persistent variant that also staged weights and predictions through a ring of shared-memory stages fed by a producer warp
//  was measured at 55-61 us against this kernel's time: shared memory caps it at ~30 pixel warps per SM and the per-warp
//  dependency chains then set the pace -- profiles/r02/blend_pipeline_experiment.patch.)
// ---------------------------------------------------------------------------------------------
#define PRV2_SEG_SPAN 512          // canvas columns staged per row and segment (floats)
#define PRV2_SEG_LIST 128          // covering patches per row and segment
#define PRV2_SEG_ROWB ((PRV2_SEG_SPAN + 4) * 4)

__device__ __forceinline__ void seg_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
template <int OFF>
__device__ __forceinline__ float seg_lds(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
  return v;
}

struct SegTables { const int* a_b; const int* c_b; const float* c_l1; };     // byte offsets 4*a_ix, 4*c_i0 per raw column; l1 per raw column

template <int MODE>
__global__ void __launch_bounds__(256, 8) blend_raw_seg_kernel(const float* __restrict__ avg_c, const float* __restrict__ cnt_c, int Hc, int Wc,
                                                            const float* __restrict__ preds, const int32_t* __restrict__ starts, int n,
                                                            int ph, int pw, int rh, int rw, int H, int W, float* __restrict__ out,
                                                            float* __restrict__ out_cnt, const float* __restrict__ num_in, RawScales sc,
                                                            SegTables tb, RawPrep prep, int segw) {
  __shared__ __align__(16) float s_rows[3][PRV2_SEG_SPAN + 4];
  __shared__ __align__(16) int4 s_ent[PRV2_SEG_LIST];             // {x0a, rw + c, moff, soff}
  __shared__ int s_poff[PRV2_SEG_LIST];
  __shared__ int s_m;
  __shared__ float s_ly1;
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x, lane = tid & 31;
  const int y = blockIdx.y;
  const int xs = blockIdx.x * segw, xe = min(xs + segw, W);
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar), rows0 = (uint32_t)__cvta_generic_to_shared(&s_rows[0][0]);
  // first staged canvas column (multiple of 4 floats), in bytes: both resamplings are monotonic in x
  const int lo4 = min(__ldg(tb.a_b + xs), __ldg(tb.c_b + xs)) & ~15;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)min(PRV2_SEG_SPAN * 4, Wc * 4 - lo4);            // Wc and lo are multiples of 4 floats
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes * 3u) : "memory");
    const BilinearTap ty = ac_tap(sc.bs_y, y, Hc);                                    // utils.py:42-43, row part
    s_ly1 = ty.l1;
    seg_bulk_g2s(rows0, reinterpret_cast<const char*>(avg_c + (size_t)nearest_src(y, sc.ns_y, Hc) * Wc) + lo4, bytes, bar);
    seg_bulk_g2s(rows0 + PRV2_SEG_ROWB, reinterpret_cast<const char*>(cnt_c + (size_t)ty.i0 * Wc) + lo4, bytes, bar);
    seg_bulk_g2s(rows0 + 2 * PRV2_SEG_ROWB, reinterpret_cast<const char*>(cnt_c + (size_t)ty.i1 * Wc) + lo4, bytes, bar);
  } else if (tid >= 32 && tid < 64) {                             // warp 1 lists the patches covering row y x [xs, xe), draw order kept
    int m = 0;
    for (int base = 0; base < n; base += 32) {
      const int k = base + lane;
      int py0 = 0, px0 = 0;
      bool hit = false;
      if (k < n) {
        py0 = starts[k * 2 + 0]; px0 = starts[k * 2 + 1];
        hit = y >= py0 && y < py0 + rh && px0 < xe && px0 + rw > xs;
      }
      const unsigned b = __ballot_sync(0xffffffffu, hit);
      if (hit) {
        const int pos = m + __popc(b & ((1u << lane) - 1));
        const int ly = y - py0, c = px0 & 3, x0a = px0 - c;
        s_ent[pos] = make_int4(x0a, rw + c, (c * rh + ly) * prep.pitch - x0a, c * prep.pitch - x0a);
        s_poff[pos] = (MODE == 2) ? 0 : (k * ph + nearest_src(ly, sc.ps_y, ph)) * pw;    // baseline_pretrain.py:210 (nearest), row part
      }
      m += __popc(b);
    }
    if (lane == 0) s_m = m;
  }
  // column-only part: shared-memory addresses of the canvas taps, 1 - l1
  const int xb = xs + tid * 4;
  const bool active = xb < xe;                                    // W % 4 == 0: an active group is a full group
  uint32_t sa[4], si[4];
  float l0[4], l1[4];
  {
    const int xl = active ? xb : xs;
    const int4 ta = __ldg(reinterpret_cast<const int4*>(tb.a_b + xl));
    const int4 ti = __ldg(reinterpret_cast<const int4*>(tb.c_b + xl));
    const float4 tl = __ldg(reinterpret_cast<const float4*>(tb.c_l1 + xl));
    const uint32_t r0 = rows0 - (uint32_t)lo4;
    sa[0] = r0 + ta.x; sa[1] = r0 + ta.y; sa[2] = r0 + ta.z; sa[3] = r0 + ta.w;
    si[0] = r0 + ti.x; si[1] = r0 + ti.y; si[2] = r0 + ti.z; si[3] = r0 + ti.w;
    l1[0] = tl.x; l1[1] = tl.y; l1[2] = tl.z; l1[3] = tl.w;
#pragma unroll
    for (int q = 0; q < 4; ++q) l0[q] = __fsub_rn(1.0f, l1[q]);
  }
  __syncthreads();                                                // list written, barrier initialised
  {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
  }
  if (lo4 + PRV2_SEG_SPAN * 4 >= Wc * 4) {
    // the segment reaches the last canvas column: tap i1 == i0 there (ac_tap), i.e. column Wc repeats column Wc - 1
    if (tid < 3) s_rows[tid][Wc - lo4 / 4] = s_rows[tid][Wc - 1 - lo4 / 4];
    __syncthreads();
  }
  if (!active) return;
  float avg[4], cnt[4], c0[4];
  {
    const float ly1 = s_ly1, ly0 = __fsub_rn(1.0f, ly1);          // (ac_tap's row weights)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      avg[q] = seg_lds<0>(sa[q]);
      const float v00 = seg_lds<PRV2_SEG_ROWB>(si[q]), v01 = seg_lds<PRV2_SEG_ROWB + 4>(si[q]);
      const float v10 = seg_lds<2 * PRV2_SEG_ROWB>(si[q]), v11 = seg_lds<2 * PRV2_SEG_ROWB + 4>(si[q]);
      const float r0 = __fmaf_rn(l0[q], v00, __fmul_rn(l1[q], v01));                  // ac_blend, ATen order
      const float r1 = __fmaf_rn(l0[q], v10, __fmul_rn(l1[q], v11));
      cnt[q] = __fmaf_rn(ly0, r0, __fmul_rn(ly1, r1));
      c0[q] = cnt[q];
    }
  }
  const int m = s_m;
  for (int i = 0; i < m; ++i) {
    const int4 e = s_ent[i];
    if ((unsigned)(xb - e.x) >= (unsigned)e.y) continue;          // group entirely outside this patch
    const float4 c4 = __ldg(reinterpret_cast<const float4*>(prep.mask4 + (e.z + xb)));
    const float ct[4] = {c4.x, c4.y, c4.z, c4.w};
    if (MODE == 2) {
#pragma unroll
      for (int q = 0; q < 4; ++q) if (ct[q] > 0.f) cnt[q] = __fadd_rn(cnt[q], ct[q]);
    } else {
      const int4 sx = __ldg(reinterpret_cast<const int4*>(prep.srci4 + (e.w + xb)));
      const int po = s_poff[i];                                   // 32-bit element offsets: one add + one IMAD.WIDE per gather
      const float p[4] = {__ldg(preds + (po + sx.x)), __ldg(preds + (po + sx.y)), __ldg(preds + (po + sx.z)), __ldg(preds + (po + sx.w))};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // utils.py:31-36 as a select, not a branch: a zero weight (the padding of an edge group) leaves the pixel untouched
        const float den = __fadd_rn(cnt[q], ct[q]);
        const float qv = __fdiv_rn(__fadd_rn(__fmul_rn(p[q], ct[q]), __fmul_rn(cnt[q], avg[q])), den);
        const bool pos = ct[q] > 0.f;
        avg[q] = pos ? qv : avg[q];
        cnt[q] = pos ? den : cnt[q];
      }
    }
  }
  const size_t o = (size_t)y * W + xb;
  if (MODE == 2) {
    const float4 nin = __ldcs(reinterpret_cast<const float4*>(num_in + o));
    const float num[4] = {nin.x, nin.y, nin.z, nin.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) avg[q] = (cnt[q] > c0[q]) ? __fdiv_rn(__fadd_rn(__fmul_rn(avg[q], c0[q]), num[q]), cnt[q]) : avg[q];
  }
  __stcs(reinterpret_cast<float4*>(out + o), make_float4(avg[0], avg[1], avg[2], avg[3]));
  if (out_cnt) __stcs(reinterpret_cast<float4*>(out_cnt + o), make_float4(cnt[0], cnt[1], cnt[2], cnt[3]));
}

// host replay of the column index math (plain IEEE fp32, same operations as nearest_src / ac_tap on the device)
static int h_nearest_src(int dst, float scale, int n_in) { const int v = (int)floorf((float)dst * scale); return v < n_in - 1 ? v : n_in - 1; }
static int h_ac_i0(float scale, int dst, int n_in) { const int v = (int)(scale * (float)dst); return v < n_in - 1 ? v : n_in - 1; }

static int g_seg_w = 0, g_seg_off = -1;            // A/B knobs (prv2_debug_blend_generic / PRV2_BLEND_SEG=0)
// bit 1: segment kernel off; bits 8-15: warps per CTA (0 = automatic)
static void seg_knobs(int on) { g_seg_off = (on & 2) ? 1 : 0; g_seg_w = (on >> 8) & 255; }
static bool seg_disabled() {
  if (g_seg_off < 0) { const char* e = getenv("PRV2_BLEND_SEG"); g_seg_off = (e && e[0] == '0') ? 1 : 0; }
  return g_seg_off == 1;
}

// Can the segment kernel take this call, and with how many warps per CTA (128*w pixels per segment)?
static bool seg_plan(int Wc, int W, int pw, int rh, int rw, int n, int ph, const RawScales& sc, const void* prep, const void* a, const void* b,
                     const void* c, const void* d, const void* e, const void* preds, int* warps) {
  if (seg_disabled() || (W & 3) || (Wc & 3) || n > PRV2_SEG_LIST || (n > 0 && !prep)) return false;
  if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d | (uintptr_t)e | (uintptr_t)preds) & 15) != 0) return false;
  const long long pitch = raw_prep_pitch(rw);
  if (4ll * rh * pitch + W >= (1ll << 29) || (long long)n * ph * pw + pw >= (1ll << 29) || (long long)Wc * 4 >= (1ll << 29)) return false;   // 32-bit byte offsets
  static const int order[] = {6, 8, 5, 7, 4, 3, 2};
  int best_w = 0; long long best_waste = -1;
  for (int k = 0; k < (int)(sizeof(order) / sizeof(order[0])); ++k) {
    const int w = g_seg_w > 0 ? g_seg_w : order[k];
    const int segw = 128 * w, nseg = cdiv(W, segw);
    bool fits = w >= 2 && w <= 8;                                 // (__launch_bounds__ of the kernel; warp 1 builds the list)
    for (int s = 0; s < nseg && fits; ++s) {
      const int xs = s * segw, xl = (xs + segw < W ? xs + segw : W) - 1;
      const int a0 = h_nearest_src(xs, sc.ns_x, Wc), i0 = h_ac_i0(sc.bs_x, xs, Wc);
      const int a1 = h_nearest_src(xl, sc.ns_x, Wc), i1 = h_ac_i0(sc.bs_x, xl, Wc) + 1;
      const int lo = (a0 < i0 ? a0 : i0) & ~3, hi = (a1 > i1 ? a1 : i1);              // hi <= Wc (column Wc = the repeated last column)
      if (hi - lo >= PRV2_SEG_SPAN) fits = false;
    }
    if (!fits) { if (g_seg_w > 0) return false; continue; }
    const long long waste = (long long)nseg * segw - W;
    if (best_waste < 0 || waste < best_waste) { best_waste = waste; best_w = w; }
    if (g_seg_w > 0 || waste == 0) break;
  }
  if (!best_w) return false;
  *warps = best_w;
  return true;
}

template <int MODE>
static bool launch_seg(const float* avg_c, const float* cnt_c, int Hc, int Wc, const float* preds, const int32_t* starts, int n, int ph, int pw, int rh,
                       int rw, int H, int W, float* out, float* out_cnt, const float* num_in, const RawScales& sc, const SegTables& tb, const RawPrep& pv,
                       int warps, cudaStream_t stream) {
  const int segw = warps * 128;
  blend_raw_seg_kernel<MODE><<<dim3(cdiv(W, segw), H), warps * 32, 0, stream>>>(avg_c, cnt_c, Hc, Wc, preds, starts, n, ph, pw, rh, rw, H, W, out, out_cnt,
                                                                                num_in, sc, tb, pv, segw);
  ++g_seg_launches;
  return true;
}

template <int MODE>
static void launch_raw(const float* avg_c, const float* cnt_c, int Hc, int Wc, const float* preds, const uint8_t* own, const int32_t* starts,
                       int n, int ph, int pw, const float* rmask, int rh, int rw, int H, int W, float* out, float* out_cnt,
                       const float* num_in, const RawScales& sc, const void* prep, cudaStream_t stream) {
  dim3 grid(1, H);
  RawTables tb;
  int warps = 0;
  if (MODE != 1 && !blend_generic_forced() && seg_plan(Wc, W, pw, rh, rw, n, ph, sc, prep, avg_c, cnt_c, out, out_cnt, num_in, preds, &warps) &&
      raw_tables_get(Wc, W, pw, rw, sc, stream, &tb)) {
    constexpr int M = MODE == 1 ? 0 : MODE;                       // (MODE 1 never gets here; keeps the instantiation list to modes 0 and 2)
    const RawPrep pv = raw_prep_view(prep, rh, rw);
    SegTables st;
    st.a_b = tb.a_b; st.c_b = tb.c_b; st.c_l1 = tb.c_l1;
    if (launch_seg<M>(avg_c, cnt_c, Hc, Wc, preds, starts, n, ph, pw, rh, rw, H, W, out, out_cnt, num_in, sc, st, pv, warps, stream))
      return;
  }
  if (!blend_generic_forced() && raw_tables_get(Wc, W, pw, rw, sc, stream, &tb)) {
    blend_raw_tab_kernel<MODE><<<grid, 256, 0, stream>>>(avg_c, cnt_c, Hc, Wc, preds, own, starts, n, ph, pw, rmask, rh, rw, H, W, out, out_cnt,
                                                         num_in, sc, tb, raw_prep_view(prep, rh, rw));
  } else {
    blend_raw_kernel<MODE><<<grid, 256, 0, stream>>>(avg_c, cnt_c, Hc, Wc, preds, own, starts, n, ph, pw, rmask, rh, rw, H, W, out, out_cnt, num_in, sc);
  }
}

static int check_raw(const char* fn, int n, int ph, int pw, int rh, int rw, int H, int W) {
  if (!(n >= 0 && n <= PRV2_MAX_RANDOM)) { set_error("%s: 0 <= n <= %d required (got %d)", fn, PRV2_MAX_RANDOM, n); return PRV2_EINVAL; }
  if (!(ph > 0 && pw > 0 && rh > 0 && rw > 0 && H > 0 && W > 0 && H <= 65535)) { set_error("%s: bad shape", fn); return PRV2_EINVAL; }
  return PRV2_OK;
}

extern "C" int prv2_blend_raw(const float* avg_c, const float* cnt_c, int Hc, int Wc, const float* preds, const int32_t* starts, int n,
                              int ph, int pw, const float* rmask, int rh, int rw, int H, int W, float* out, float* out_cnt,
                              const void* prep, prv2_stream_t stream) {
  PRV2_CHECK_ARG(avg_c && cnt_c && out, "prv2_blend_raw: null pointer");
  PRV2_CHECK_ARG(n == 0 || (preds && starts && rmask), "prv2_blend_raw: null patch inputs");
  int rc = check_raw("prv2_blend_raw", n, ph, pw, rh, rw, H, W);
  if (rc) return rc;
  launch_raw<0>(avg_c, cnt_c, Hc, Wc, preds, nullptr, starts, n, ph, pw, rmask, rh, rw, H, W, out, out_cnt, nullptr,
                raw_scales(Hc, Wc, H, W, ph, pw, rh, rw), n ? prep : nullptr, (cudaStream_t)stream);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_blend_partial_raw(const float* preds, const uint8_t* own, const int32_t* starts, int n, int ph, int pw,
                                      const float* rmask, int rh, int rw, int H, int W, float* num_r, const void* prep, prv2_stream_t stream) {
  PRV2_CHECK_ARG(preds && own && starts && rmask && num_r, "prv2_blend_partial_raw: null pointer");
  int rc = check_raw("prv2_blend_partial_raw", n, ph, pw, rh, rw, H, W);
  if (rc) return rc;
  if (n == 0) return PRV2_OK;
  launch_raw<1>(nullptr, nullptr, 1, 1, preds, own, starts, n, ph, pw, rmask, rh, rw, H, W, num_r, nullptr, nullptr,
                raw_scales(1, 1, H, W, ph, pw, rh, rw), prep, (cudaStream_t)stream);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_blend_finalize_raw(const float* avg_c, const float* cnt_c, int Hc, int Wc, const float* num_r, const int32_t* starts,
                                       int n, const float* rmask, int rh, int rw, int H, int W, float* out, float* out_cnt,
                                       const void* prep, prv2_stream_t stream) {
  PRV2_CHECK_ARG(avg_c && cnt_c && num_r && out, "prv2_blend_finalize_raw: null pointer");
  PRV2_CHECK_ARG(n == 0 || (starts && rmask), "prv2_blend_finalize_raw: null patch inputs");
  int rc = check_raw("prv2_blend_finalize_raw", n, 1, 1, rh, rw, H, W);
  if (rc) return rc;
  launch_raw<2>(avg_c, cnt_c, Hc, Wc, nullptr, nullptr, starts, n, 1, 1, rmask, rh, rw, H, W, out, out_cnt, num_r,
                raw_scales(Hc, Wc, H, W, 1, 1, rh, rw), n ? prep : nullptr, (cudaStream_t)stream);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}
