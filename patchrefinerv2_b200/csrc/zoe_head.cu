// ZoeDepth metric-bins head, the per-pixel part (external/zoedepth/models/zoedepth/zoedepth_v1.py:173-219): everything between
// the head's 1x1 convolutions (those run on prv2_umma_gemm) is memory-bound arithmetic over 64 bin centres per pixel, kept in fp32.
//   prv2_zoe_attractor          AttractorLayerUnnormed (layers/attractor.py:139-208, inv_attractor :45-57): bilinear(align_corners)
//                               resampling of the previous level's centres fused with the attractor update
//   prv2_zoe_logbinomial_depth  ConditionalLogBinomial's tail + LogBinomial + the expectation (layers/dist_layers.py:29-122,
//                               zoedepth_v1.py:212-219): softplus -> (p, T) -> 64-way log-binomial softmax -> sum_k prob_k * centre_k
// Bin centres are channels-last [B, h, w, K] so that consecutive threads read consecutive bins.
#include "common.cuh"

using namespace prv2;

namespace {

__device__ __forceinline__ float softplus1(float x) { return x > 20.f ? x : log1pf(expf(x)); }      // F.softplus(beta=1, threshold=20)

// block = 256 threads = 4 pixels x 64 bins (K <= 64 per pass; larger K loops)
__global__ void __launch_bounds__(256) zoe_attractor_kernel(const float* __restrict__ a_raw, int a_ld, int na, const float* __restrict__ b_prev, int hp,
                                                            int wp, int prev_is_raw, float* __restrict__ b_out, long long pixels, int h, int w, int K,
                                                            float alpha, int mean) {
  __shared__ float s_a[4][32];
  const int slot = threadIdx.x >> 6, kk = threadIdx.x & 63;
  const long long pix = (long long)blockIdx.x * 4 + slot;
  const bool live = pix < pixels;
  if (live && kk < na) s_a[slot][kk] = softplus1(a_raw[pix * a_ld + kk]);          // A = softplus(net(x))  (attractor.py:181-184)
  __syncthreads();
  if (!live) return;
  const int x = (int)(pix % w), y = (int)((pix / w) % h);
  const long long n = pix / ((long long)w * h);
  const BilinearTap ty = ac_tap(ac_scale(hp, h), y, hp), tx = ac_tap(ac_scale(wp, w), x, wp);
  const float* base = b_prev + (size_t)n * hp * wp * K;
  for (int k = kk; k < K; k += 64) {
    float v00 = base[((size_t)ty.i0 * wp + tx.i0) * K + k], v01 = base[((size_t)ty.i0 * wp + tx.i1) * K + k];
    float v10 = base[((size_t)ty.i1 * wp + tx.i0) * K + k], v11 = base[((size_t)ty.i1 * wp + tx.i1) * K + k];
    if (prev_is_raw) { v00 = softplus1(v00); v01 = softplus1(v01); v10 = softplus1(v10); v11 = softplus1(v11); }   // seed centres = softplus(regressor) (localbins_layers.py:95)
    const float bc = ac_blend(ty, tx, v00, v01, v10, v11);                        // b_centers = interpolate(b_prev, align_corners=True) (:186-187)
    float acc = 0.f;
    for (int i = 0; i < na; ++i) {
      const float dx = __fsub_rn(s_a[slot][i], bc);
      acc = __fadd_rn(acc, __fdiv_rn(dx, __fadd_rn(1.0f, __fmul_rn(alpha, __fmul_rn(dx, dx)))));     // dx / (1 + alpha * dx^2)  (:45-57, gamma = 2)
    }
    if (mean) acc = __fdiv_rn(acc, (float)na);
    b_out[(size_t)pix * K + k] = __fadd_rn(bc, acc);                              // b_new = b_centers + delta (:205)
  }
}

// warp per pixel, lane owns bins lane, lane + 32, ...
__global__ void __launch_bounds__(256) zoe_depth_kernel(const float* __restrict__ pt_raw, int pt_ld, const float* __restrict__ centers, int hb, int wb,
                                                        float* __restrict__ depth, long long pixels, int H, int W, int K, float min_temp,
                                                        float max_temp) {
  const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pix >= pixels) return;
  const float* pt = pt_raw + pix * pt_ld;
  const float p_eps = 1e-4f, eps = 1e-4f;
  const float p0 = softplus1(pt[0]) + p_eps, p1 = softplus1(pt[1]) + p_eps;      // dist_layers.py:95-109
  const float t0 = softplus1(pt[2]) + p_eps, t1 = softplus1(pt[3]) + p_eps;
  const float prob = p0 / (p0 + p1);
  const float temp = (max_temp - min_temp) * (t0 / (t0 + t1)) + min_temp;
  const float one_minus = fminf(fmaxf(1.f - prob, eps), 1.f), xk = fminf(fmaxf(prob, eps), 1.f);     // LogBinomial.forward (:52-69)
  const float lx = logf(xk), l1 = logf(one_minus);
  const float nb = (float)(K - 1) + 1e-7f;                                        // log_binom(n, k) (:29-33), n = K - 1
  const float nlogn = nb * logf(nb);
  const int x = (int)(pix % W), y = (int)((pix / W) % H);
  const long long n = pix / ((long long)W * H);
  const BilinearTap ty = ac_tap(ac_scale(hb, H), y, hb), tx = ac_tap(ac_scale(wb, W), x, wb);
  const float* base = centers + (size_t)n * hb * wb * K;
  float yk[8], ck[8];                                                             // K <= 256
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = lane + 32 * j;
    yk[j] = -INFINITY; ck[j] = 0.f;
    if (k < K) {
      const float kf = (float)k + 1e-7f;
      const float lb = nlogn - kf * logf(kf) - (nb - kf) * logf(nb - kf + 1e-7f);
      yk[j] = (lb + (float)k * lx + (float)(K - 1 - k) * l1) / temp;
      m = fmaxf(m, yk[j]);
      ck[j] = ac_blend(ty, tx, base[((size_t)ty.i0 * wb + tx.i0) * K + k], base[((size_t)ty.i0 * wb + tx.i1) * K + k],
                       base[((size_t)ty.i1 * wb + tx.i0) * K + k], base[((size_t)ty.i1 * wb + tx.i1) * K + k]);     // zoedepth_v1.py:217-218
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float se = 0.f, sc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (lane + 32 * j < K) {
      const float e = expf(yk[j] - m);
      se += e;
      sc = fmaf(e, ck[j], sc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { se += __shfl_xor_sync(0xffffffffu, se, o); sc += __shfl_xor_sync(0xffffffffu, sc, o); }
  if (lane == 0) depth[pix] = sc / se;                                            // sum_k softmax(y / T)_k * centre_k (:219)
}

}  // namespace

extern "C" int prv2_zoe_attractor(const float* a_raw, int a_ld, int n_attractors, const float* b_prev, int hp, int wp, int prev_is_raw, float* b_out, int B,
                                  int h, int w, int n_bins, float alpha, int mean, prv2_stream_t stream) {
  PRV2_CHECK_ARG(a_raw && b_prev && b_out, "prv2_zoe_attractor: null pointer");
  PRV2_CHECK_ARG(B > 0 && h > 0 && w > 0 && hp > 0 && wp > 0 && n_bins > 0 && n_attractors > 0 && n_attractors <= 32 && a_ld >= n_attractors,
                 "prv2_zoe_attractor: bad shape (1 <= n_attractors <= 32)");
  const long long pixels = (long long)B * h * w;
  zoe_attractor_kernel<<<(unsigned)((pixels + 3) / 4), 256, 0, (cudaStream_t)stream>>>(a_raw, a_ld, n_attractors, b_prev, hp, wp, prev_is_raw, b_out, pixels, h, w,
                                                                                       n_bins, alpha, mean);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_zoe_logbinomial_depth(const float* pt_raw, int pt_ld, const float* centers, int hb, int wb, float* depth, int B, int H, int W, int n_bins,
                                          float min_temp, float max_temp, prv2_stream_t stream) {
  PRV2_CHECK_ARG(pt_raw && centers && depth, "prv2_zoe_logbinomial_depth: null pointer");
  PRV2_CHECK_ARG(B > 0 && H > 0 && W > 0 && hb > 0 && wb > 0 && n_bins > 1 && n_bins <= 256 && pt_ld >= 4, "prv2_zoe_logbinomial_depth: bad shape (2 <= n_bins <= 256)");
  const long long pixels = (long long)B * H * W;
  zoe_depth_kernel<<<(unsigned)((pixels + 7) / 8), 256, 0, (cudaStream_t)stream>>>(pt_raw, pt_ld, centers, hb, wb, depth, pixels, H, W, n_bins, min_temp, max_temp);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}
