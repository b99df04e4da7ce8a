// Hand-written memory-bound kernels around the tcgen05 contractions: LayerNorm, patch-embed
// im2col, token assembly, bilinear resampling, depth-channel injection, the Cout=1 final conv,
// layout converters.  Activations are channels-last bf16 (hi[,lo]) so that a warp always touches
// contiguous channels; every thread moves 16-byte vectors.
#include "common.cuh"

using namespace prv2;

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct alignas(8) bf16x4 { bf16 v[4]; };
__device__ __forceinline__ void act_store4(bf16* hi, bf16* lo, size_t i, const float (&in)[4]) {
  if (lo) {                                   // (hi, lo) fp16 pair, see common.cuh
    __half2 h2[2], l2[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float a = sat16(in[2 * k]), b = sat16(in[2 * k + 1]);
      h2[k] = __floats2half2_rn(a, b);
      const float2 f = __half22float2(h2[k]);
      l2[k] = __floats2half2_rn(a - f.x, b - f.y);
    }
    *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<const uint2*>(h2);
    *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<const uint2*>(l2);
  } else {
    bf16x4 a;
#pragma unroll
    for (int k = 0; k < 4; ++k) a.v[k] = f2bf(in[k]);
    *reinterpret_cast<bf16x4*>(hi + i) = a;
  }
}

// ------------------------------------------------------------------ LayerNorm (block.py:56,68)
// warp per row, row kept in registers (D <= 1024), two-pass mean / variance.
template <bool X3, bool GELU>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int rows, int D, const float* __restrict__ w,
                                                        const float* __restrict__ b, float eps, int drop_period, bf16* __restrict__ oh,
                                                        bf16* __restrict__ ol_, int out_cs) {
  bf16* const ol = X3 ? ol_ : nullptr;
  constexpr bool gelu = GELU;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  long long orow = row;
  if (drop_period > 0) {
    if (row % drop_period == 0) return;                  // class token dropped (dinov2.py:311)
    orow = row - row / drop_period - 1;
  }
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  const int nv = D >> 7;                                 // float4 chunks per lane
  float4 v[8];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) { v[j] = xr[j * 32 + lane]; sum += v[j].x + v[j].y + v[j].z + v[j].w; }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const float a = v[j].x - mean, bb = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      sq += a * a + bb * bb + c * c + d * d;
    }
  const float rstd = 1.0f / sqrtf(warp_sum(sq) / (float)D + eps);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const int c0 = (j * 32 + lane) * 4;
      const float4 ww = *reinterpret_cast<const float4*>(w + c0), bb = *reinterpret_cast<const float4*>(b + c0);
      float o[4] = {(v[j].x - mean) * rstd * ww.x + bb.x, (v[j].y - mean) * rstd * ww.y + bb.y, (v[j].z - mean) * rstd * ww.z + bb.z,
                    (v[j].w - mean) * rstd * ww.w + bb.w};
      if (gelu) { o[0] = gelu_erf(o[0]); o[1] = gelu_erf(o[1]); o[2] = gelu_erf(o[2]); o[3] = gelu_erf(o[3]); }
      act_store4(oh, ol, (size_t)orow * out_cs + c0, o);
    }
}

// ------------------------------------------------- patch-embed im2col (patch_embed.py:69-82)
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ crops, int B, int H, int W, bf16* __restrict__ oh,
                                                       bf16* __restrict__ ol, int Kp, long long total) {
  const int gw = W / 14, gh = H / 14, groups = Kp / 8;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx % groups);
    const long long row = idx / groups;
    const int tx = (int)(row % gw), ty = (int)((row / gw) % gh), b = (int)(row / ((long long)gw * gh));
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int col = g * 8 + k;
      if (col < 588) {
        const int c = col / 196, r = col % 196, ky = r / 14, kx = r % 14;
        const float px = __ldg(crops + (((size_t)b * 3 + c) * H + (ty * 14 + ky)) * W + tx * 14 + kx);
        o[k] = (px - mean[c]) / stdv[c];               // dpt.py:183
      } else {
        o[k] = 0.f;
      }
    }
    act_store8(oh, ol, (size_t)row * Kp + g * 8, o);
  }
}

__global__ void __launch_bounds__(256) assemble_tokens_kernel(const float* __restrict__ emb, const float* __restrict__ cls,
                                                              const float* __restrict__ pos, int B, int T, int D, float* __restrict__ x,
                                                              long long total) {
  const int dv = D / 4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int d4 = (int)(idx % dv) * 4;
    const long long r = idx / dv;
    const int t = (int)(r % (T + 1)), b = (int)(r / (T + 1));
    const float4 a = (t == 0) ? *reinterpret_cast<const float4*>(cls + d4)
                              : *reinterpret_cast<const float4*>(emb + ((size_t)b * T + (t - 1)) * D + d4);
    const float4 pp = *reinterpret_cast<const float4*>(pos + (size_t)t * D + d4);
    *reinterpret_cast<float4*>(x + (size_t)r * D + d4) = make_float4(a.x + pp.x, a.y + pp.y, a.z + pp.z, a.w + pp.w);
  }
}

// --------------------------------------- bilinear align_corners=True on channels-last acts
// grid (ceil(OW*C/8 / 256), ceil(OH / RS_ROWS), N).  A thread owns 8 channels of one output column and walks RS_ROWS
// consecutive output rows: the horizontal blend of a source row (2 x 16-byte loads, bf16 unpack, 8 FMAs) is computed once
// and reused by every output row that needs it (an up-sampling ratio r reuses each source row ~r times).  The first
// version recomputed all four taps per output and was ALU-pipe bound (77 % ALU, 2.0 TB/s) on the unpack / index work.
constexpr int RS_ROWS = 8;

// X3 / RELU are compile-time: with a runtime `lo` pointer the (hi, lo) plane code was predicated off in bf16 mode but still
// issued -- 645 of the kernel's ~2700 SASS instructions -- in a kernel that is ALU-pipe bound.
template <bool X3, bool RELU>
__global__ void __launch_bounds__(256) resize_act_kernel(const bf16* __restrict__ ih, const bf16* __restrict__ il_, int h, int w, int C,
                                                         int in_cs, bf16* __restrict__ oh, bf16* __restrict__ ol_, int OH, int OW,
                                                         int out_cs, float sy, float sx) {
  const bf16* const il = X3 ? il_ : nullptr;
  bf16* const ol = X3 ? ol_ : nullptr;
  constexpr bool relu = RELU;
  const unsigned cv = (unsigned)C >> 3, total = (unsigned)OW * cv;
  const unsigned idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx / cv), c8 = (int)(idx % cv) * 8;
  const int y0 = blockIdx.y * RS_ROWS, n = blockIdx.z;
  const BilinearTap tx = ac_tap(sx, x, w);
  const size_t ibase = (size_t)n * h * w;
  auto hblend = [&](int row, float (&o)[8]) {
    float a[8], b[8];
    act_load8(ih, il, (ibase + (size_t)row * w + tx.i0) * in_cs + c8, a);
    act_load8(ih, il, (ibase + (size_t)row * w + tx.i1) * in_cs + c8, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = __fmaf_rn(tx.l0, a[k], __fmul_rn(tx.l1, b[k]));
  };
  int r_lo = -1, r_hi = -1;
  float hb_lo[8], hb_hi[8];
#pragma unroll
  for (int yy = 0; yy < RS_ROWS; ++yy) {
    const int y = y0 + yy;
    if (y >= OH) break;
    const BilinearTap ty = ac_tap(sy, y, h);
    if (ty.i0 != r_lo) {
      if (ty.i0 == r_hi) {
#pragma unroll
        for (int k = 0; k < 8; ++k) hb_lo[k] = hb_hi[k];
      } else {
        hblend(ty.i0, hb_lo);
      }
      r_lo = ty.i0;
    }
    if (ty.i1 != r_hi) {
      if (ty.i1 == r_lo) {
#pragma unroll
        for (int k = 0; k < 8; ++k) hb_hi[k] = hb_lo[k];
      } else {
        hblend(ty.i1, hb_hi);
      }
      r_hi = ty.i1;
    }
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      o[k] = __fmaf_rn(ty.l0, hb_lo[k], __fmul_rn(ty.l1, hb_hi[k]));
      if (relu) o[k] = fmaxf(o[k], 0.f);
    }
    act_store8(oh, ol, (((size_t)n * OH + y) * OW + x) * out_cs + c8, o);
  }
}

// -------------------------- depth maps as an im2col source (fusion_model.py:94-96, 16-17)
// The reference concatenates the two depth maps (bilinear-resized to the level) to a feature tensor that then goes
// through a 3x3 conv.  Two extra channels per tap would cost one mostly empty 64-wide K chunk per tap (9 per conv);
// instead the 3x3 neighbourhood of both maps is laid out ONCE per level as 18 channels
//   out[n, y, x, (r*3+s)*2 + d] = resized(pred_d)[y+r-1, x+s-1]   (0 outside the map = the conv's zero padding)
// and enters every conv of the level as a single 1x1 K segment.  Channels 18..23 are zero.  grid (ceil(OW/256), OH, N)
__global__ void __launch_bounds__(256) depth_taps_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int H, int W,
                                                         bf16* __restrict__ oh, bf16* __restrict__ ol, int OH, int OW, int out_cs,
                                                         float sy, float sx) {
  const int x = blockIdx.x * 256 + threadIdx.x;
  if (x >= OW) return;
  const int y = blockIdx.y, n = blockIdx.z;
  const size_t b = (size_t)n * H * W;
  float o[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) o[i] = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int yy = y + r - 1;
    if (yy < 0 || yy >= OH) continue;
    const BilinearTap ty = ac_tap(sy, yy, H);
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int xx = x + s - 1;
      if (xx < 0 || xx >= OW) continue;
      const BilinearTap tx = ac_tap(sx, xx, W);
      const size_t i00 = b + (size_t)ty.i0 * W + tx.i0, i01 = b + (size_t)ty.i0 * W + tx.i1, i10 = b + (size_t)ty.i1 * W + tx.i0,
                   i11 = b + (size_t)ty.i1 * W + tx.i1;
      o[(r * 3 + s) * 2 + 0] = ac_blend(ty, tx, __ldg(p1 + i00), __ldg(p1 + i01), __ldg(p1 + i10), __ldg(p1 + i11));
      o[(r * 3 + s) * 2 + 1] = ac_blend(ty, tx, __ldg(p2 + i00), __ldg(p2 + i01), __ldg(p2 + i10), __ldg(p2 + i11));
    }
  }
  const size_t ob = (((size_t)n * OH + y) * OW + x) * out_cs;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    float t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = o[g * 8 + k];
    act_store8(oh, ol, ob + g * 8, t);
  }
}

// ------------------------------ final 3x3 conv to one channel + base + clamp (fusion_model.py:113-118)
// The channel contraction runs on the tensor cores as a 1x1 conv producing the 9 tap responses
// Y[p, t] = sum_c feat[p, c] * w[t, c] (fp32, 16 floats per pixel); this stencil gathers
// out[p] = clamp(base[p] + sum_t Y[p + off_t, t], 0) with zero padding.  grid (ceil(W/256), H, N).
__global__ void __launch_bounds__(256) tap_stencil_kernel(const float* __restrict__ Y, int H, int W, int ld,
                                                          const float* __restrict__ base, float* __restrict__ out) {
  const int x = blockIdx.x * 256 + threadIdx.x;
  if (x >= W) return;
  const int y = blockIdx.y, n = blockIdx.z;
  float acc = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int yy = y + r - 1;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int xx = x + s - 1;
      if (xx < 0 || xx >= W) continue;
      acc += __ldg(Y + (((size_t)n * H + yy) * W + xx) * ld + r * 3 + s);
    }
  }
  const size_t o = ((size_t)n * H + y) * W + x;
  if (base) acc = fmaxf(base[o] + acc, 0.f);
  out[o] = acc;
}

// ------------------------------ final 3x3 conv to ONE channel + base + clamp, fused (fusion_model.py:113-118)
// HBM-bound: the C-channel feature map is read once.  CTA = 30x30 output pixels = a 32x32 halo tile.  Phase 1: every halo
// pixel gets its 9 tap responses Y[t] = sum_c feat[c] * w[t][c] (fp32) into shared memory; out-of-image pixels are the
// conv's zero padding.  FOUR lanes share a pixel (lane `sub` owns the 8-channel groups sub, sub+4, ...), so a warp-wide
// 16-byte load touches 8 cache lines instead of 32 (thread-per-pixel loads were L1-tag bound), and a thread works on four
// pixels at a time so one weight vector from shared memory feeds all four (the weight reads were shared-memory-bandwidth bound at two); the four partial sums meet in two shuffle steps.
// Phase 2: out = clamp(base + sum_{r,s} Y[(y+r-1, x+s-1)][r*3+s], 0).  Replaces a 9-column GEMM (tile overheads on a
// K=128, N=16 problem left it at ~1 TB/s) plus the tap-stencil pass over its 64-byte-per-pixel output.
#define FC_T 30
#define FC_HALO (FC_T + 2)
#define FC_PIX (FC_HALO * FC_HALO)
#define FC_NPX 4                 // pixels a thread works on at once: one shared-memory weight vector feeds all of them
template <bool X3>
__global__ void __launch_bounds__(256) final_conv_kernel(const bf16* __restrict__ fh, const bf16* __restrict__ fl, int H, int W, int C, int cs,
                                                         const float* __restrict__ wt, const float* __restrict__ base, float* __restrict__ out) {
  extern __shared__ float fc_smem[];
  float* const s_w = fc_smem;                 // [9][C]
  float* const s_y = fc_smem + 9 * C;         // [9][FC_PIX]
  const int n = blockIdx.z, y0 = blockIdx.y * FC_T, x0 = blockIdx.x * FC_T;
  for (int i = threadIdx.x; i < 9 * C; i += 256) s_w[i] = wt[i];
  __syncthreads();
  const size_t img = (size_t)n * H * W;
  const int sub = threadIdx.x & 3, grp = threadIdx.x >> 2;            // 64 pixel slots per pass, FC_NPX pixels per thread and iteration
  for (int pbase = 0; pbase < FC_PIX; pbase += 64 * FC_NPX) {
    int pi[FC_NPX];
    const bf16* ph[FC_NPX];
    const bf16* pl[FC_NPX];
    bool ok[FC_NPX];
#pragma unroll
    for (int k = 0; k < FC_NPX; ++k) {
      pi[k] = pbase + k * 64 + grp;
      const int hy = pi[k] / FC_HALO, hx = pi[k] - hy * FC_HALO;
      const int yy = y0 + hy - 1, xx = x0 + hx - 1;
      ok[k] = yy >= 0 && yy < H && xx >= 0 && xx < W;
      const size_t off = ok[k] ? (img + (size_t)yy * W + xx) * cs : 0;
      ph[k] = fh + off;
      pl[k] = X3 ? fl + off : nullptr;
    }
    float acc[FC_NPX][9];
#pragma unroll
    for (int k = 0; k < FC_NPX; ++k)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[k][t] = 0.f;
    for (int c0 = sub * 8; c0 < C; c0 += 32) {
      float v[FC_NPX][8];
#pragma unroll
      for (int k = 0; k < FC_NPX; ++k) {
        if (ok[k]) {
          act_load8(ph[k], pl[k], c0, v[k]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
        }
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 w0 = *reinterpret_cast<const float4*>(s_w + t * C + c0), w1 = *reinterpret_cast<const float4*>(s_w + t * C + c0 + 4);
#pragma unroll
        for (int k = 0; k < FC_NPX; ++k) {
          float a = acc[k][t];
          a = fmaf(v[k][0], w0.x, a); a = fmaf(v[k][1], w0.y, a); a = fmaf(v[k][2], w0.z, a); a = fmaf(v[k][3], w0.w, a);
          a = fmaf(v[k][4], w1.x, a); a = fmaf(v[k][5], w1.y, a); a = fmaf(v[k][6], w1.z, a); a = fmaf(v[k][7], w1.w, a);
          acc[k][t] = a;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < FC_NPX; ++k)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float a = acc[k][t];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        acc[k][t] = a;
      }
    // lane `sub` stores taps sub, sub+4, sub+8 of both pixels
#pragma unroll
    for (int k = 0; k < FC_NPX; ++k)
#pragma unroll
      for (int t = 0; t < 9; ++t)
        if ((t & 3) == sub) s_y[t * FC_PIX + pi[k]] = acc[k][t];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < FC_T * FC_T; i += 256) {
    const int ty = i / FC_T, tx = i - ty * FC_T;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) a += s_y[(r * 3 + q) * FC_PIX + (ty + r) * FC_HALO + tx + q];
    const size_t o = img + (size_t)y * W + x;
    if (base) a = fmaxf(base[o] + a, 0.f);
    out[o] = a;
  }
}

__global__ void __launch_bounds__(256) nchw_to_act_kernel(const float* __restrict__ in, int N, int C, int H, int W, bf16* __restrict__ oh,
                                                          bf16* __restrict__ ol, int out_cs, long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % out_cs);
    const long long pix = idx / out_cs;
    const int hw = (int)(pix % ((long long)H * W));
    const int n = (int)(pix / ((long long)H * W));
    const float v = c < C ? in[((size_t)n * C + c) * H * W + hw] : 0.f;
    act_store(oh, ol, (size_t)idx, v);
  }
}

__global__ void __launch_bounds__(256) act_to_nchw_kernel(const bf16* __restrict__ ih, const bf16* __restrict__ il, int N, int C, int H,
                                                          int W, int in_cs, float* __restrict__ out, long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int hw = (int)(idx % ((long long)H * W));
    const int c = (int)((idx / ((long long)H * W)) % C);
    const int n = (int)(idx / ((long long)H * W * C));
    out[idx] = act_load(ih, il, ((size_t)n * H * W + hw) * in_cs + c);
  }
}

// out[(py*2+px)*N + n, y2, x2, :] = in[n, 2*y2 + py, 2*x2 + px, :], zero where that source pixel does not exist: with odd H or W
// the four phases of a stride-2 conv have ceil / floor sizes, all are stored at the conv's output size ceil(H/2) x ceil(W/2)
// (the missing row / column is exactly the conv's zero padding, dpt.py:72-80).
__global__ void __launch_bounds__(256) phase_split_kernel(const bf16* __restrict__ ih, const bf16* __restrict__ il, int N, int H, int W,
                                                          int C, int in_cs, bf16* __restrict__ oh, bf16* __restrict__ ol, int out_cs,
                                                          long long total) {
  const int cv = C / 8, H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const size_t plane = (size_t)N * H2 * W2 * out_cs;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % cv) * 8;
    long long pix = idx / cv;
    const int x2 = (int)(pix % W2);
    pix /= W2;
    const int y2 = (int)(pix % H2);
    pix /= H2;
    const int n = (int)(pix % N);
    const int ph = (int)(pix / N);
    const int y = 2 * y2 + (ph >> 1), x = 2 * x2 + (ph & 1);
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (y < H && x < W) act_load8(ih, il, (((size_t)n * H + y) * W + x) * in_cs + c8, v);
    act_store8(oh, ol ? ol : nullptr, ph * plane + (((size_t)n * H2 + y2) * W2 + x2) * out_cs + c8, v);
  }
}

__global__ void __launch_bounds__(256) split_f32_kernel(const float* __restrict__ in, long long n, bf16* __restrict__ hi, bf16* __restrict__ lo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    act_store(hi, lo, (size_t)i, in[i]);
}

// Depthwise k x k convolution (k = 3 or 5, stride 1 or 2, symmetric zero padding k / 2) on channels-last acts, BatchNorm already
// folded into (w, bias) by the host, optional ReLU: the dw_start / dw_mid stages of MobileNetV4's UniversalInvertedResidual blocks
// (the V2 family's light-weight refiner encoder, lightweight_refiner.py:259-262).  Thread = one output pixel x 8 channels: k*k
// 16-byte activation vectors (consecutive threads walk the channel groups of a pixel, so a warp reads whole 128-byte lines) and
// k*k*8 fp32 weights from the [k*k, C] table (L1-resident), fp32 accumulation in tap order (ky outer, kx inner).  HBM-bound.
template <int K>
__global__ void __launch_bounds__(256) dwconv_kernel(const bf16* __restrict__ ih, const bf16* __restrict__ il, int N, int H, int W, int C, int in_cs,
                                                     const float* __restrict__ w, const float* __restrict__ bias, int stride, int relu,
                                                     bf16* __restrict__ oh, bf16* __restrict__ ol, int OH, int OW, int out_cs, long long total) {
  const int cv = C / 8;
  constexpr int P = K / 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % cv) * 8;
    long long pix = idx / cv;
    const int ox = (int)(pix % OW);
    pix /= OW;
    const int oy = (int)(pix % OH);
    const int n = (int)(pix / OH);
    float acc[8];
    {
      const float4 b0 = bias ? __ldg(reinterpret_cast<const float4*>(bias + c8)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 b1 = bias ? __ldg(reinterpret_cast<const float4*>(bias + c8 + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int y = oy * stride + ky - P;
      if (y < 0 || y >= H) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int x = ox * stride + kx - P;
        if (x < 0 || x >= W) continue;
        float v[8];
        act_load8(ih, il, (((size_t)n * H + y) * W + x) * in_cs + c8, v);
        const float* wk = w + (size_t)(ky * K + kx) * C + c8;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wk)), w1 = __ldg(reinterpret_cast<const float4*>(wk + 4));
        acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]); acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
        acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]); acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
      }
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
    }
    act_store8(oh, ol, (((size_t)n * OH + oy) * OW + ox) * out_cs + c8, acc);
  }
}

// Encoder input of the light-weight refiner (lightweight_refiner.py:293-298): (crop - mean) / std for the three colour channels,
// the coarse depth as a fourth channel (coarse_condition), channels 4..7 zero -> one 16-byte channels-last vector per pixel.
__global__ void __launch_bounds__(256) encoder_input_kernel(const float* __restrict__ crops, const float* __restrict__ depth, int N, int H, int W,
                                                            float m0, float m1, float m2, float s0, float s1, float s2, bf16* __restrict__ oh,
                                                            bf16* __restrict__ ol, int out_cs, long long total) {
  const size_t hw = (size_t)H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const size_t n = (size_t)(idx / hw), px = (size_t)(idx % hw);
    const float* c = crops + n * 3 * hw + px;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    v[0] = __fdiv_rn(__fsub_rn(c[0], m0), s0);
    v[1] = __fdiv_rn(__fsub_rn(c[hw], m1), s1);
    v[2] = __fdiv_rn(__fsub_rn(c[2 * hw], m2), s2);
    if (depth) v[3] = depth[n * hw + px];
    act_store8(oh, ol, (size_t)idx * out_cs, v);
  }
}

int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" int prv2_layernorm(const float* x, int rows, int D, const float* w, const float* b, float eps, int drop_period,
                              prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs, prv2_stream_t stream) {
  PRV2_CHECK_ARG(x && w && b && out_hi, "prv2_layernorm: null pointer");
  PRV2_CHECK_ARG(rows >= 0 && D > 0 && D % 128 == 0 && D <= 1024 && out_cs % 4 == 0 && out_cs >= D, "prv2_layernorm: D must be a multiple of 128, <= 1024 (got %d)", D);
  if (rows == 0) return PRV2_OK;
  if (out_lo) layernorm_kernel<true, false><<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, D, w, b, eps, drop_period, (bf16*)out_hi, (bf16*)out_lo, out_cs);
  else layernorm_kernel<false, false><<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, D, w, b, eps, drop_period, (bf16*)out_hi, nullptr, out_cs);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_layernorm_gelu(const float* x, int rows, int D, const float* w, const float* b, float eps,
                                   prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs, prv2_stream_t stream) {
  PRV2_CHECK_ARG(x && w && b && out_hi, "prv2_layernorm_gelu: null pointer");
  PRV2_CHECK_ARG(rows >= 0 && D > 0 && D % 128 == 0 && D <= 1024 && out_cs % 4 == 0 && out_cs >= D, "prv2_layernorm_gelu: D must be a multiple of 128, <= 1024 (got %d)", D);
  if (rows == 0) return PRV2_OK;
  if (out_lo) layernorm_kernel<true, true><<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, D, w, b, eps, 0, (bf16*)out_hi, (bf16*)out_lo, out_cs);
  else layernorm_kernel<false, true><<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(x, rows, D, w, b, eps, 0, (bf16*)out_hi, nullptr, out_cs);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_patchify(const float* crops, int B, int H, int W, prv2_bf16* out_hi, prv2_bf16* out_lo, int Kp, prv2_stream_t stream) {
  PRV2_CHECK_ARG(crops && out_hi, "prv2_patchify: null pointer");
  PRV2_CHECK_ARG(B >= 0 && H > 0 && W > 0 && H % 14 == 0 && W % 14 == 0 && Kp >= 588 && Kp % 8 == 0, "prv2_patchify: H,W must be multiples of 14; Kp>=588, Kp%%8==0");
  if (B == 0) return PRV2_OK;
  const long long total = (long long)B * (H / 14) * (W / 14) * (Kp / 8);
  patchify_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(crops, B, H, W, (bf16*)out_hi, (bf16*)out_lo, Kp, total);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_assemble_tokens(const float* emb, const float* cls, const float* pos, int B, int T, int D, float* x, prv2_stream_t stream) {
  PRV2_CHECK_ARG(emb && cls && pos && x, "prv2_assemble_tokens: null pointer");
  PRV2_CHECK_ARG(B >= 0 && T > 0 && D > 0 && D % 4 == 0, "prv2_assemble_tokens: bad shape");
  if (B == 0) return PRV2_OK;
  const long long total = (long long)B * (T + 1) * (D / 4);
  assemble_tokens_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(emb, cls, pos, B, T, D, x, total);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_resize_bilinear_act(const prv2_bf16* in_hi, const prv2_bf16* in_lo, int N, int h, int w, int C, int in_cs,
                                        prv2_bf16* out_hi, prv2_bf16* out_lo, int oh, int ow, int out_cs, int relu, prv2_stream_t stream) {
  PRV2_CHECK_ARG(in_hi && out_hi, "prv2_resize_bilinear_act: null pointer");
  PRV2_CHECK_ARG(N >= 0 && h > 0 && w > 0 && oh > 0 && ow > 0 && C > 0 && C % 8 == 0 && in_cs % 8 == 0 && out_cs % 8 == 0, "prv2_resize_bilinear_act: bad shape");
  PRV2_CHECK_ARG(oh <= 65535 && N <= 65535, "prv2_resize_bilinear_act: grid too large");
  if (N == 0) return PRV2_OK;
  const float sy = oh > 1 ? (float)(h - 1) / (float)(oh - 1) : 0.f, sx = ow > 1 ? (float)(w - 1) / (float)(ow - 1) : 0.f;
  PRV2_CHECK_ARG((long long)ow * (C / 8) < (1LL << 31), "prv2_resize_bilinear_act: row too long");
  dim3 grid(cdiv((long long)ow * (C / 8), 256), cdiv(oh, RS_ROWS), N);
  PRV2_CHECK_ARG((in_lo == nullptr) == (out_lo == nullptr), "prv2_resize_bilinear_act: lo planes must both be present or absent");
#define PRV2_RS(X, R) resize_act_kernel<X, R><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)in_hi, (const bf16*)in_lo, h, w, C, in_cs, \
                                                                                       (bf16*)out_hi, (bf16*)out_lo, oh, ow, out_cs, sy, sx)
  if (in_lo) { if (relu) PRV2_RS(true, true); else PRV2_RS(true, false); }
  else { if (relu) PRV2_RS(false, true); else PRV2_RS(false, false); }
#undef PRV2_RS
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_depth_taps(const float* pred1, const float* pred2, int N, int H, int W, prv2_bf16* out_hi, prv2_bf16* out_lo, int oh,
                               int ow, int out_cs, prv2_stream_t stream) {
  PRV2_CHECK_ARG(pred1 && pred2 && out_hi, "prv2_depth_taps: null pointer");
  PRV2_CHECK_ARG(N >= 0 && H > 0 && W > 0 && oh > 0 && ow > 0 && out_cs % 8 == 0 && out_cs >= 24, "prv2_depth_taps: pitch must be a multiple of 8, >= 24");
  PRV2_CHECK_ARG(oh <= 65535 && N <= 65535, "prv2_depth_taps: grid too large");
  if (N == 0) return PRV2_OK;
  const float sy = oh > 1 ? (float)(H - 1) / (float)(oh - 1) : 0.f, sx = ow > 1 ? (float)(W - 1) / (float)(ow - 1) : 0.f;
  dim3 grid(cdiv(ow, 256), oh, N);
  depth_taps_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred1, pred2, H, W, (bf16*)out_hi, (bf16*)out_lo, oh, ow, out_cs, sy, sx);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_tap_stencil(const float* taps, int N, int H, int W, int ld, const float* base, float* out, prv2_stream_t stream) {
  PRV2_CHECK_ARG(taps && out, "prv2_tap_stencil: null pointer");
  PRV2_CHECK_ARG(N >= 0 && H > 0 && W > 0 && ld >= 9 && H <= 65535 && N <= 65535, "prv2_tap_stencil: bad shape");
  if (N == 0) return PRV2_OK;
  dim3 grid(cdiv(W, 256), H, N);
  tap_stencil_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(taps, H, W, ld, base, out);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_final_conv3x3(const prv2_bf16* feat_hi, const prv2_bf16* feat_lo, int N, int H, int W, int C, int cs, const float* w9c,
                                  const float* base, float* out, prv2_stream_t stream) {
  PRV2_CHECK_ARG(feat_hi && w9c && out, "prv2_final_conv3x3: null pointer");
  PRV2_CHECK_ARG(N >= 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && cs % 8 == 0 && cs >= C && N <= 65535, "prv2_final_conv3x3: C and pitch must be multiples of 8");
  const size_t smem = (size_t)(9 * C + 9 * FC_PIX) * sizeof(float);
  PRV2_CHECK_ARG(smem <= 200 * 1024, "prv2_final_conv3x3: C=%d too wide for the shared-memory weight table", C);
  if (N == 0) return PRV2_OK;
  dim3 grid(cdiv(W, FC_T), cdiv(H, FC_T), N);
  if (feat_lo) {
    PRV2_CUDA(cudaFuncSetAttribute(final_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    final_conv_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)feat_hi, (const bf16*)feat_lo, H, W, C, cs, w9c, base, out);
  } else {
    PRV2_CUDA(cudaFuncSetAttribute(final_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    final_conv_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)feat_hi, nullptr, H, W, C, cs, w9c, base, out);
  }
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_nchw_f32_to_act(const float* in, int N, int C, int H, int W, prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs,
                                    prv2_stream_t stream) {
  PRV2_CHECK_ARG(in && out_hi && out_cs >= C, "prv2_nchw_f32_to_act: bad arguments");
  const long long total = (long long)N * H * W * out_cs;
  if (total == 0) return PRV2_OK;
  nchw_to_act_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, N, C, H, W, (bf16*)out_hi, (bf16*)out_lo, out_cs, total);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_act_to_nchw_f32(const prv2_bf16* in_hi, const prv2_bf16* in_lo, int N, int C, int H, int W, int in_cs, float* out,
                                    prv2_stream_t stream) {
  PRV2_CHECK_ARG(in_hi && out && in_cs >= C, "prv2_act_to_nchw_f32: bad arguments");
  const long long total = (long long)N * C * H * W;
  if (total == 0) return PRV2_OK;
  act_to_nchw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)in_hi, (const bf16*)in_lo, N, C, H, W, in_cs, out, total);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_phase_split(const prv2_bf16* in_hi, const prv2_bf16* in_lo, int N, int H, int W, int C, int in_cs, prv2_bf16* out_hi,
                                prv2_bf16* out_lo, int out_cs, prv2_stream_t stream) {
  PRV2_CHECK_ARG(in_hi && out_hi, "prv2_phase_split: null pointer");
  PRV2_CHECK_ARG(H > 0 && W > 0 && C % 8 == 0 && in_cs % 8 == 0 && out_cs % 8 == 0, "prv2_phase_split: C and the pitches must be multiples of 8");
  PRV2_CHECK_ARG((in_lo == nullptr) == (out_lo == nullptr), "prv2_phase_split: lo planes must both be present or absent");
  const long long total = 4LL * N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  if (total == 0) return PRV2_OK;
  phase_split_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)in_hi, (const bf16*)in_lo, N, H, W, C, in_cs,
                                                                           (bf16*)out_hi, (bf16*)out_lo, out_cs, total);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_split_f32(const float* in, int64_t n, prv2_bf16* hi, prv2_bf16* lo, prv2_stream_t stream) {
  PRV2_CHECK_ARG(in && hi && n >= 0, "prv2_split_f32: bad arguments");
  if (n == 0) return PRV2_OK;
  split_f32_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, n, (bf16*)hi, (bf16*)lo);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_dwconv(const prv2_bf16* in_hi, const prv2_bf16* in_lo, int N, int H, int W, int C, int in_cs, const float* w, const float* bias,
                           int k, int stride, int relu, prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs, prv2_stream_t stream) {
  PRV2_CHECK_ARG(in_hi && out_hi && w, "prv2_dwconv: null pointer");
  PRV2_CHECK_ARG((k == 3 || k == 5) && (stride == 1 || stride == 2), "prv2_dwconv: k must be 3 or 5 and stride 1 or 2 (got k=%d stride=%d)", k, stride);
  PRV2_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && in_cs % 8 == 0 && out_cs % 8 == 0 && in_cs >= C && out_cs >= C,
                 "prv2_dwconv: C and the pitches must be multiples of 8");
  PRV2_CHECK_ARG((in_lo == nullptr) == (out_lo == nullptr), "prv2_dwconv: lo planes must both be present or absent");
  PRV2_CHECK_ARG(((uintptr_t)w & 15) == 0 && ((uintptr_t)bias & 15) == 0, "prv2_dwconv: weights / bias must be 16-byte aligned");
  const int OH = (H + 2 * (k / 2) - k) / stride + 1, OW = (W + 2 * (k / 2) - k) / stride + 1;
  const long long total = (long long)N * OH * OW * (C / 8);
  if (k == 3)
    dwconv_kernel<3><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)in_hi, (const bf16*)in_lo, N, H, W, C, in_cs, w, bias, stride,
                                                                            relu, (bf16*)out_hi, (bf16*)out_lo, OH, OW, out_cs, total);
  else
    dwconv_kernel<5><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)in_hi, (const bf16*)in_lo, N, H, W, C, in_cs, w, bias, stride,
                                                                            relu, (bf16*)out_hi, (bf16*)out_lo, OH, OW, out_cs, total);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}

extern "C" int prv2_encoder_input(const float* crops, const float* depth, int N, int H, int W, const float* mean3, const float* std3,
                                  prv2_bf16* out_hi, prv2_bf16* out_lo, int out_cs, prv2_stream_t stream) {
  PRV2_CHECK_ARG(crops && out_hi && mean3 && std3, "prv2_encoder_input: null pointer");
  PRV2_CHECK_ARG(N > 0 && H > 0 && W > 0 && out_cs >= 8 && out_cs % 8 == 0, "prv2_encoder_input: bad shape");
  const long long total = (long long)N * H * W;
  encoder_input_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(crops, depth, N, H, W, mean3[0], mean3[1], mean3[2], std3[0], std3[1],
                                                                              std3[2], (bf16*)out_hi, (bf16*)out_lo, out_cs, total);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}
