// Persistent TMA + tcgen05 implicit-GEMM for sm_100a: every dense contraction of the per-patch
// network (ViT linears, patch-embed, DPT convs / deconvs, FusionUnet convs) runs through this
// one kernel.
//
//   D[m, n] = sum over K-segments s, channels c:  A_s[img, h + dh_s, w + dw_s, c] * Wt[n, k(s, c)]
//
// * M tile = 128 output pixels = a (tile_h x tile_w) spatial block of ONE image, fetched by a 4-D
//   TMA box {64 ch, tile_w, tile_h, 1}; out-of-bounds box elements are zero-filled by TMA, which
//   implements the conv padding, ragged edges and K tails for free.  A linear layer is the case
//   H = 1, W = rows, tile = 128 x 1.
// * The box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle = the canonical
//   K-major UMMA operand layout, so the same bytes feed tcgen05.mma through a shared-memory
//   descriptor; weights [Cout_pad, Ktot] (K-major) are fetched with a 2-D box {64, block_n}.
// * Accumulators live in TMEM (2 x block_n fp32 columns, double buffered) so the epilogue of tile
//   i overlaps the MMAs of tile i+1.  Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer
//   (one elected lane each), warps 2..5 = epilogue (one TMEM lane quadrant each).
// * Epilogues fuse bias, ReLU / erf-GELU, channels-first LayerNorm+GELU, residual adds,
//   LayerScale*gamma + fp32 residual stream, the ConvTranspose pixel shuffle and the DPT
//   sigmoid head.
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

using namespace prv2;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                        // bf16 per K chunk (128 B rows)
constexpr int A_STAGE_BYTES = BM * BK * 2;    // 16 KB
constexpr int NUM_THREADS = 192;
constexpr int EPI_THREADS = 128;
constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 220 * 1024;       // data stages; barriers sit behind
constexpr uint64_t SPIN_LIMIT_NS = 4000000000ull;   // a wedged pipeline traps instead of hanging the box

struct alignas(64) KParams {
  CUtensorMap tmA[PRV2_MAX_SRC];
  CUtensorMap tmB;
  int16_t seg_src[PRV2_MAX_SEG];
  int16_t seg_dh[PRV2_MAX_SEG];
  int16_t seg_dw[PRV2_MAX_SEG];
  uint8_t seg_chunks[PRV2_MAX_SEG];           // 64-wide chunks in this segment
  uint8_t seg_last[PRV2_MAX_SEG];             // 16-wide MMA slices in the last chunk (1..4)
  int32_t n_seg;
  int32_t N, H, W, Cout;
  int32_t tile_w, tile_h, tiles_w, tiles_h, tiles_n, total_tiles;
  int32_t block_n, stages, b_stage_bytes, tmem_cols;
  int32_t epi, act;
  const float* bias;
  const float* gamma;
  const float* beta;
  float eps, head_scale;
  bf16* out_hi; bf16* out_lo; int32_t out_cs;
  bf16* relu_hi; bf16* relu_lo; int32_t relu_cs;
  const bf16* res_hi; const bf16* res_lo; int32_t res_cs;
  const bf16* res2_hi; const bf16* res2_lo; int32_t res2_cs;
  float* out_f32; int32_t out_f32_ld;
  int32_t shuffle_k;
  int32_t row_map_period, row_map_extra, row_map_offset;
};
static_assert(sizeof(KParams) <= 4096, "kernel parameter block too large");

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 0x3fff) == 0 && globaltimer_ns() - t0 > SPIN_LIMIT_NS) {
      printf("prv2_umma_gemm: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128-byte-swizzled operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(1)<<16
// | SBO(1024 B>>4)<<32 | version(1)<<46 | layout SWIZZLE_128B(2)<<61.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// 16 consecutive channels of an act tensor
__device__ __forceinline__ void load16(const bf16* hi, const bf16* lo, size_t i, float (&v)[16]) {
  float a[8], b[8];
  act_load8(hi, lo, i, a);
  act_load8(hi, lo, i + 8, b);
#pragma unroll
  for (int k = 0; k < 8; ++k) { v[k] = a[k]; v[8 + k] = b[k]; }
}
__device__ __forceinline__ void store_n(bf16* hi, bf16* lo, size_t i, const float (&v)[16], int n_valid, bool relu) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    if (g * 8 + 8 <= n_valid) {
      float t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = relu ? fmaxf(v[g * 8 + k], 0.f) : v[g * 8 + k];
      act_store8(hi, lo, i + g * 8, t);
    }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1) umma_gemm_kernel(const __grid_constant__ KParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stages = p.stages;
  const uint32_t stage_bytes = A_STAGE_BYTES + p.b_stage_bytes;
  const uint32_t bar_base = smem_base + stages * stage_bytes;
  // barrier map: full[s] | empty[s] | tmem_full[2] | tmem_empty[2] | tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_THREADS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < PRV2_MAX_SRC; ++s) prefetch_tmap(&p.tmA[s]);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int tiles_per_img = p.tiles_h * p.tiles_w;

  if (warp == 0) {
    // ================================ TMA producer =========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.tiles_n, mt = tile / p.tiles_n;
        const int img = mt / tiles_per_img, r = mt % tiles_per_img;
        const int h0 = (r / p.tiles_w) * p.tile_h, w0 = (r % p.tiles_w) * p.tile_w;
        const int n0 = nt * p.block_n;
        int kcol = 0;
        for (int s = 0; s < p.n_seg; ++s) {
          const CUtensorMap* map = &p.tmA[p.seg_src[s]];
          const int dh = p.seg_dh[s], dw = p.seg_dw[s];
          const int chunks = p.seg_chunks[s];
          for (int c = 0; c < chunks; ++c) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t a_dst = smem_base + stage * stage_bytes;
            mbar_expect_tx(full_bar(stage), stage_bytes);
            tma_load_4d(a_dst, map, full_bar(stage), c * BK, w0 + dw, h0 + dh, img);
            tma_load_2d(a_dst + A_STAGE_BYTES, &p.tmB, full_bar(stage), kcol, n0);
            kcol += BK;
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===========================================
    if (lane == 0) {
      // cute::UMMA::InstrDescriptor: c=F32(1)<<4 | a=BF16(1)<<7 | b=BF16(1)<<10 | N>>3 <<17 | M>>4 <<24, both K-major
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * p.block_n;
        uint32_t accumulate = 0;
        for (int s = 0; s < p.n_seg; ++s) {
          const int chunks = p.seg_chunks[s];
          for (int c = 0; c < chunks; ++c) {
            const int slices = (c == chunks - 1) ? p.seg_last[s] : 4;
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = smem_base + stage * stage_bytes;
            const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + A_STAGE_BYTES);
            for (int k = 0; k < slices; ++k) {
              // +32 bytes per 16-element K slice inside the 128-byte swizzle row (encoded >>4)
              tc_mma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, accumulate);
              accumulate = 1;
            }
            tc_commit(empty_bar(stage));
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
        tc_commit(tfull_bar(acc));
      }
    }
  } else {
    // ================================ epilogue =============================================
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    const int n_chunks = p.block_n >> 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int nt = tile % p.tiles_n, mt = tile / p.tiles_n;
      const int img = mt / tiles_per_img, r = mt % tiles_per_img;
      const int h = (r / p.tiles_w) * p.tile_h + row / p.tile_w;
      const int w = (r % p.tiles_w) * p.tile_w + row % p.tile_w;
      const bool valid = (h < p.H) && (w < p.W);
      const int n0 = nt * p.block_n;
      const size_t m = ((size_t)img * p.H + h) * p.W + w;
      size_t orow = m;
      if (p.row_map_period > 0) orow = m + (m / p.row_map_period) * p.row_map_extra + p.row_map_offset;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * p.block_n;
      float v[16];

      if (p.epi == PRV2_EPI_LN_GELU) {
        // channels-first LayerNorm over the Cout channels of this pixel (two-pass, as convs.py:24-27)
        float sum = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
          tc_ld16(taddr + c * 16, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) if (n0 + c * 16 + j < p.Cout) sum += v[j];
        }
        const float mean = sum / (float)p.Cout;
        float sq = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
          tc_ld16(taddr + c * 16, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) if (n0 + c * 16 + j < p.Cout) { const float d = v[j] - mean; sq += d * d; }
        }
        const float rstd = 1.0f / sqrtf(sq / (float)p.Cout + p.eps);
        for (int c = 0; c < n_chunks; ++c) {
          const int n = n0 + c * 16;
          tc_ld16(taddr + c * 16, v);
          if (valid && n < p.Cout) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int nn = min(n + j, p.Cout - 1);
              v[j] = gelu_erf(__ldg(p.gamma + nn) * ((v[j] - mean) * rstd) + __ldg(p.beta + nn));
            }
            store_n(p.out_hi, p.out_lo, orow * p.out_cs + n, v, min(16, p.Cout - n), false);
          }
        }
      } else if (p.epi == PRV2_EPI_HEAD) {
        float dot = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
          const int n = n0 + c * 16;
          tc_ld16(taddr + c * 16, v);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n + j < p.Cout) dot += fmaxf(v[j] + __ldg(p.bias + n + j), 0.f) * __ldg(p.gamma + n + j);
        }
        if (valid) p.out_f32[orow] = p.head_scale / (1.0f + expf(-(dot + __ldg(p.beta))));
      } else {
        for (int c = 0; c < n_chunks; ++c) {
          const int n = n0 + c * 16;
          tc_ld16(taddr + c * 16, v);
          if (!valid || n >= p.Cout) continue;
          const int nv = min(16, p.Cout - n);
          if (p.epi == PRV2_EPI_SHUFFLE) {
            // n = (ky*k + kx)*Cout_real + co ; Cout here counts k*k*Cout_real columns
            const int k = p.shuffle_k, co_n = p.Cout / (k * k);
            const int tap = n / co_n, co = n % co_n, ky = tap / k, kx = tap % k;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += p.bias ? __ldg(p.bias + co + j) : 0.f;
            const size_t opix = ((size_t)img * (p.H * k) + (h * k + ky)) * (p.W * k) + (w * k + kx);
            store_n(p.out_hi, p.out_lo, opix * p.out_cs + co, v, 16, false);
            continue;
          }
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __ldg(p.bias + min(n + j, p.Cout - 1));
          }
          if (p.epi == PRV2_EPI_RESID_F32) {
            float* x = p.out_f32 + orow * p.out_f32_ld + n;
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j < nv) x[j] = x[j] + __ldg(p.gamma + n + j) * v[j];
            continue;
          }
          if (p.epi == PRV2_EPI_F32) {
            float* x = p.out_f32 + orow * p.out_f32_ld + n;
#pragma unroll
            for (int j = 0; j < 16; ++j) if (j < nv) x[j] = v[j];
            continue;
          }
          if (p.act == PRV2_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          } else if (p.act == PRV2_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
          }
          if (p.res_hi) {
            float rr[16];
            if (nv == 16) load16(p.res_hi, p.res_lo, m * p.res_cs + n, rr);
            else { for (int j = 0; j < 16; ++j) rr[j] = j < nv ? act_load(p.res_hi, p.res_lo, m * p.res_cs + n + j) : 0.f; }
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += rr[j];
          }
          if (p.res2_hi) {
            float rr[16];
            if (nv == 16) load16(p.res2_hi, p.res2_lo, m * p.res2_cs + n, rr);
            else { for (int j = 0; j < 16; ++j) rr[j] = j < nv ? act_load(p.res2_hi, p.res2_lo, m * p.res2_cs + n + j) : 0.f; }
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += rr[j];
          }
          if (p.out_hi) store_n(p.out_hi, p.out_lo, orow * p.out_cs + n, v, nv, false);
          if (p.relu_hi) store_n(p.relu_hi, p.relu_lo, orow * p.relu_cs + n, v, nv, true);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------- host side
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)ptr;
  }
  return fn;
}

int g_num_sms = 0;

}  // namespace

extern "C" int prv2_umma_gemm(const prv2_gemm_desc* d, prv2_stream_t stream) {
  PRV2_CHECK_ARG(d != nullptr, "prv2_umma_gemm: null desc");
  PRV2_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->Cout > 0, "prv2_umma_gemm: bad output shape");
  PRV2_CHECK_ARG(d->tile_w > 0 && d->tile_h > 0 && d->tile_w * d->tile_h == BM && d->tile_w <= 256 && d->tile_h <= 256,
                 "prv2_umma_gemm: tile_w*tile_h must be 128 (got %dx%d)", d->tile_w, d->tile_h);
  PRV2_CHECK_ARG(d->block_n >= 16 && d->block_n <= 256 && d->block_n % 16 == 0, "prv2_umma_gemm: block_n %d not in 16..256 step 16", d->block_n);
  PRV2_CHECK_ARG(d->n_src >= 1 && d->n_src <= PRV2_MAX_SRC && d->n_seg >= 1 && d->n_seg <= PRV2_MAX_SEG, "prv2_umma_gemm: bad n_src/n_seg");
  PRV2_CHECK_ARG(d->weight && d->Cout_pad % d->block_n == 0 && d->Cout_pad >= d->Cout && d->Ktot % BK == 0,
                 "prv2_umma_gemm: weight must be [Cout_pad (multiple of block_n), Ktot (multiple of 64)]");
  PRV2_CHECK_ARG(((uintptr_t)d->weight & 15) == 0, "prv2_umma_gemm: weight not 16-byte aligned");
  auto enc = get_encode();
  if (!enc) { set_error("prv2_umma_gemm: cuTensorMapEncodeTiled unavailable"); return PRV2_ECUDA; }

  KParams p;
  memset(&p, 0, sizeof(p));
  for (int s = 0; s < d->n_src; ++s) {
    const prv2_src& src = d->src[s];
    PRV2_CHECK_ARG(src.ptr && src.C > 0 && src.cs >= src.C && src.cs % 8 == 0 && ((uintptr_t)src.ptr & 15) == 0,
                   "prv2_umma_gemm: source %d needs C>0, pitch multiple of 8, 16-byte aligned", s);
    cuuint64_t dims[4] = {(cuuint64_t)src.C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)src.cs * 2, (cuuint64_t)d->W * src.cs * 2, (cuuint64_t)d->H * d->W * src.cs * 2};
    cuuint32_t box[4] = {BK, (cuuint32_t)d->tile_w, (cuuint32_t)d->tile_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.tmA[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)src.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("prv2_umma_gemm: cuTensorMapEncodeTiled(A%d) failed (%d) C=%d cs=%d N=%d H=%d W=%d", s, (int)r, src.C, src.cs, d->N, d->H, d->W); return PRV2_ECUDA; }
  }
  for (int s = d->n_src; s < PRV2_MAX_SRC; ++s) p.tmA[s] = p.tmA[0];
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)d->Ktot * 2};
    cuuint32_t box[2] = {BK, (cuuint32_t)d->block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)d->weight, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("prv2_umma_gemm: cuTensorMapEncodeTiled(B) failed (%d) Ktot=%d Cout_pad=%d", (int)r, d->Ktot, d->Cout_pad); return PRV2_ECUDA; }
  }
  int ktot = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    const prv2_seg& sg = d->seg[s];
    PRV2_CHECK_ARG(sg.src >= 0 && sg.src < d->n_src, "prv2_umma_gemm: segment %d references source %d", s, sg.src);
    const int C = d->src[sg.src].C;
    const int chunks = (C + BK - 1) / BK;
    PRV2_CHECK_ARG(chunks <= 255, "prv2_umma_gemm: segment too long");
    p.seg_src[s] = sg.src; p.seg_dh[s] = sg.dh; p.seg_dw[s] = sg.dw;
    p.seg_chunks[s] = (uint8_t)chunks;
    const int tail = C - (chunks - 1) * BK;
    p.seg_last[s] = (uint8_t)((tail + 15) / 16);
    ktot += chunks * BK;
  }
  PRV2_CHECK_ARG(ktot == d->Ktot, "prv2_umma_gemm: Ktot %d does not match the segment table (%d)", d->Ktot, ktot);
  p.n_seg = d->n_seg;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cout = d->Cout;
  p.tile_w = d->tile_w; p.tile_h = d->tile_h;
  p.tiles_w = cdiv(d->W, d->tile_w); p.tiles_h = cdiv(d->H, d->tile_h);
  p.tiles_n = d->Cout_pad / d->block_n;
  const long long total = (long long)d->N * p.tiles_h * p.tiles_w * p.tiles_n;
  PRV2_CHECK_ARG(total < (1LL << 31), "prv2_umma_gemm: too many tiles");
  p.total_tiles = (int)total;
  p.block_n = d->block_n;
  p.b_stage_bytes = d->block_n * BK * 2;
  p.stages = SMEM_BUDGET / (A_STAGE_BYTES + p.b_stage_bytes);
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  int cols = 32;
  while (cols < 2 * d->block_n) cols <<= 1;
  p.tmem_cols = cols;
  p.epi = d->epi; p.act = d->act;
  p.bias = d->bias; p.gamma = d->gamma; p.beta = d->beta; p.eps = d->eps; p.head_scale = d->head_scale;
  p.out_hi = (bf16*)d->out_hi; p.out_lo = (bf16*)d->out_lo; p.out_cs = d->out_cs;
  p.relu_hi = (bf16*)d->relu_hi; p.relu_lo = (bf16*)d->relu_lo; p.relu_cs = d->relu_cs;
  p.res_hi = (const bf16*)d->res_hi; p.res_lo = (const bf16*)d->res_lo; p.res_cs = d->res_cs;
  p.res2_hi = (const bf16*)d->res2_hi; p.res2_lo = (const bf16*)d->res2_lo; p.res2_cs = d->res2_cs;
  p.out_f32 = d->out_f32; p.out_f32_ld = d->out_f32_ld;
  p.shuffle_k = d->shuffle_k;
  p.row_map_period = d->row_map_period; p.row_map_extra = d->row_map_extra; p.row_map_offset = d->row_map_offset;

  switch (d->epi) {
    case PRV2_EPI_STORE:
      PRV2_CHECK_ARG((d->out_hi || d->relu_hi) && d->Cout % 8 == 0, "prv2_umma_gemm: STORE needs an output and Cout%%8==0");
      PRV2_CHECK_ARG((!d->out_hi || d->out_cs % 8 == 0) && (!d->relu_hi || d->relu_cs % 8 == 0) && (!d->res_hi || d->res_cs % 8 == 0) &&
                         (!d->res2_hi || d->res2_cs % 8 == 0), "prv2_umma_gemm: channel pitches must be multiples of 8");
      break;
    case PRV2_EPI_LN_GELU:
      PRV2_CHECK_ARG(d->out_hi && d->gamma && d->beta && p.tiles_n == 1 && d->Cout % 8 == 0 && d->out_cs % 8 == 0,
                     "prv2_umma_gemm: LN_GELU needs gamma/beta, Cout<=block_n, Cout%%8==0");
      break;
    case PRV2_EPI_RESID_F32:
      PRV2_CHECK_ARG(d->out_f32 && d->gamma && d->out_f32_ld >= d->Cout, "prv2_umma_gemm: RESID_F32 needs out_f32 and gamma");
      break;
    case PRV2_EPI_F32:
      PRV2_CHECK_ARG(d->out_f32 && d->out_f32_ld >= d->Cout, "prv2_umma_gemm: F32 needs out_f32");
      break;
    case PRV2_EPI_SHUFFLE: {
      const int kk = d->shuffle_k * d->shuffle_k;
      PRV2_CHECK_ARG(d->shuffle_k >= 1 && d->out_hi && d->Cout % kk == 0 && (d->Cout / kk) % 16 == 0 && d->out_cs % 8 == 0,
                     "prv2_umma_gemm: SHUFFLE needs Cout = k*k*C with C%%16==0");
      break;
    }
    case PRV2_EPI_HEAD:
      PRV2_CHECK_ARG(d->out_f32 && d->bias && d->gamma && d->beta && p.tiles_n == 1, "prv2_umma_gemm: HEAD needs bias/w2/b2 and one N tile");
      break;
    default:
      set_error("prv2_umma_gemm: unknown epilogue %d", d->epi);
      return PRV2_EINVAL;
  }

  if (g_num_sms == 0) {
    int dev = 0;
    PRV2_CUDA(cudaGetDevice(&dev));
    PRV2_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    PRV2_CUDA(cudaFuncSetAttribute(umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET + 1024 + 256));
  }
  // always request the full budget: guarantees one CTA per SM, so a 512-column TMEM allocation can never deadlock
  const int smem = SMEM_BUDGET + 1024 + 256;
  const int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  umma_gemm_kernel<<<grid, NUM_THREADS, smem, (cudaStream_t)stream>>>(p);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}
