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
//   (one elected lane each), warps 2..9 = epilogue (two per TMEM lane quadrant); the epilogue
//   transposes 32x16 fp32 tiles through shared memory so that global traffic is 16-byte vectors
//   over whole 32-byte sectors.
// * Epilogues fuse bias, ReLU / erf-GELU, channels-first LayerNorm+GELU, residual adds,
//   LayerScale*gamma + fp32 residual stream, the ConvTranspose pixel shuffle and the DPT
//   sigmoid head.
// * CG = 2 (large problems): two CTAs of a cluster (one SM pair) run ONE tcgen05.mma.cta_group::2 of
//   256 x N: each CTA stages its own 128-pixel A tile and HALF of the weight tile, so shared-memory
//   fill traffic per FLOP drops by a third (48 KB -> 32 KB per 128x256x64 MACs) -- the kernel is
//   bound by L2 -> SM bandwidth, not by the tensor pipe.  The leader CTA issues the MMAs and
//   multicasts its commits to both CTAs' barriers; each CTA drains its own 128 TMEM lanes.
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

using namespace prv2;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                        // bf16 per K chunk (128 B rows)
constexpr int A_STAGE_BYTES = BM * BK * 2;    // 16 KB
constexpr int NUM_EPI_WARPS = 8;              // two per TMEM lane quadrant, each takes every other 16-column chunk
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int EPI_THREADS = 32 * NUM_EPI_WARPS;
constexpr int MAX_STAGES = 8;
constexpr int STG_LD = 20;                    // floats per staged row (16 + 4 pad: conflict-free float4 access both ways)
constexpr int ROW_PITCH = 144;                // row epilogue: 64 bf16 (128 B) + 16 B pad per staged row (conflict-free 16-byte stores)
constexpr int STG_BYTES_PER_WARP = 32 * ROW_PITCH;          // >= 32x16 fp32 transpose tile + LN statistics of the generic epilogue (3328 B)
constexpr int STG_BYTES_PER_WARP_RESID = 2 * STG_BYTES_PER_WARP;   // row epilogues: two 128-byte rows (fp32 residual) or one 256-byte row (bf16) per lane
constexpr int ROW_PITCH2 = 272;               // 128 bf16 (256 B) + 16 B pad
static_assert(32 * ROW_PITCH2 <= STG_BYTES_PER_WARP_RESID, "staging");
constexpr int PAR_BYTES = 2 * 3 * 256 * 4;    // per-tile bias / gamma / beta, double buffered by accumulator parity
constexpr int LN_BYTES = NUM_EPI_WARPS * 2 * 64 * 4;        // row epilogue: LayerNorm partials exchanged between the two warps of a quadrant
constexpr int SMEM_BUDGET = 179 * 1024;       // operand stages; epilogue staging, parameters and barriers sit behind
constexpr int SMEM_TOTAL = SMEM_BUDGET + NUM_EPI_WARPS * STG_BYTES_PER_WARP + PAR_BYTES + LN_BYTES + 1024 + 256;
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");
constexpr long long SPIN_LIMIT_CLK = 8000000000ll;     // ~4 s of SM clocks: a deadlock traps instead of hanging the GPU

struct alignas(64) KParams {
  CUtensorMap tmA[PRV2_MAX_SRC];
  CUtensorMap tmB;
  CUtensorMap tmOut;                          // row epilogue in TMA-store mode: the output tensor, box = one warp's 32 rows x 128 bytes
  int16_t seg_src[PRV2_MAX_SEG];
  int16_t seg_dh[PRV2_MAX_SEG];
  int16_t seg_dw[PRV2_MAX_SEG];
  uint8_t seg_chunks[PRV2_MAX_SEG];           // 64-wide chunks in this segment
  uint8_t seg_last[PRV2_MAX_SEG];             // 16-wide MMA slices in the last chunk (1..4)
  uint8_t seg_taps[PRV2_MAX_SEG];             // 1 = single tap, 3 = vertical tap group sharing one halo fetch
  int32_t n_seg;
  int32_t N, H, W, Cout;
  int32_t tile_w, tile_h, tile_w_log2, tiles_w, tiles_h, tiles_n, total_tiles;   // total_tiles counts tile PAIRS when CG == 2
  int32_t block_n, stages, b_stage_bytes, tmem_cols;   // b_stage_bytes: ONE 64-wide K block of this CTA's weight share
  int32_t kb;                                          // 64-wide K blocks per pipeline stage (2 when the N tile is narrow)
  int32_t stage_bytes, b_off, halo_bytes, halo_step;   // stage layout: [A region b_off bytes][B blocks]; halo tile bytes; bytes per dh step
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
  float acc_scale;          // accumulators are multiplied by this on their way out of TMEM (weights pre-scaled by its inverse, see the header)
  int32_t f16;              // operand planes are FP16 (the (hi, lo) pair format) instead of bf16
  int32_t stg_warp_bytes;   // epilogue staging per warp (row epilogue of the fp32 residual stream double-buffers its row)
  int32_t tma_store;        // row epilogue (bf16 outputs): 1 = ship each warp's 32 x 128-byte panels with TMA tensor stores, 0 = one bulk copy per lane
  int32_t debug;      // diagnostics (PRV2_GEMM_DEBUG): bit0 = no TMA after the first ring pass, bit1 = epilogue drains TMEM only
};
static_assert(sizeof(KParams) <= 4096, "kernel parameter block too large");

// Diagnostics (compiled in only with -DPRV2_GEMM_TRACE_BUILD): clock64 stamps of CTA 0's roles per tile, [role][tile][point], read
// back with prv2_debug_gemm_trace.  role 0 = epilogue warp 2, role 1 = MMA warp, role 2 = TMA producer.
#ifdef PRV2_GEMM_TRACE_BUILD
__device__ unsigned long long g_gemm_trace[3 * 64 * 8];
#define GTRACE(role, it, k)                                                                                  \
  do {                                                                                                       \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (it) < 64) g_gemm_trace[((role) * 64 + (it)) * 8 + (k)] = clock64(); \
  } while (0)
#else
#define GTRACE(role, it, k) do { } while (0)
#endif

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
  // try_wait suspends the thread in hardware for a while by itself; the watchdog below costs nothing until a wait has spun 4096
  // times and then only reads the SM clock (round 1 read %globaltimer in front of every contended wait)
  if (mbar_try(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 0xfff) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > SPIN_LIMIT_CLK) {
        if ((threadIdx.x & 31) == 0)
          printf("prv2_umma_gemm: mbarrier wait timed out (block %d,%d,%d warp %d, barrier @%u parity %u)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x >> 5,
                 bar & 0x3ffu, parity);
        __trap();
      }
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
// ---- cta_group::2 (CTA pair) variants
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {      // same offset in CTA `rank` of the cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Accumulator hand-back of the epilogue warps.  RELAXED: what has to be ordered before the MMA warp overwrites the accumulator are this
// warp's TMEM loads, and those are complete (tcgen05.wait::ld) and fenced (tcgen05.fence::before_thread_sync) -- no memory of the
// generic proxy is published through this barrier.  The .release.cluster form waited for the warp's outstanding global / async
// writes every tile: ~3000 clocks per tile in the per-tile traces (scripts/gemm_trace.py), a third of the epilogue loop.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {          // arrives on `bar` at the same offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One lane of a CONVERGED warp.  The single-thread roles run their loops with the whole warp converged and elect the
// issuing lane per step: under `if (lane == 0)` ptxas cannot prove uniformity and wraps every UTCHMMA / UTMALDG /
// UTCBAR in an ELECT + BRA.U.ANY loop, which made MMA issue (not the tensor pipe) the bottleneck.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
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
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16], float scale) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) * scale;
}

// K-major, 128-byte-swizzled operand descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(1)<<16
// | SBO(1024 B>>4)<<32 | version(1)<<46 | layout SWIZZLE_128B(2)<<61.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7): one MUFU.RCP + one MUFU.EX2 + 6 FMA instead of
// erff's ~25-instruction branchy path; exact-GELU error stays at fp32 rounding level.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * exp2f(-z * z * 1.4426950408889634f);
  return 0.5f * x * (1.0f + copysignf(e, x));
}

// tanh-form GELU on MUFU.TANH (5 FP32 ops + 1 MUFU): |gelu_tanh - gelu_erf| <= ~1e-3 abs, below the bf16 rounding of
// the stored activation; used by the one-pass bf16 mode only.  The erf form above costs ~17 FP32 ops + 2 MUFU per
// element and made the fc1 / FusionUnet GELU epilogues FMA-pipe bound.
__device__ __forceinline__ float gelu_tanh(float x) {
  const float x2 = x * x;
  const float inner = x * fmaf(0.0356774081f, x2, 0.7978845608f);      // sqrt(2/pi) * (x + 0.044715 x^3)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
template <int ACT>
__device__ __forceinline__ float act_fn(float x) {
  if (ACT == PRV2_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == PRV2_ACT_GELU) return gelu_fast(x);
  if (ACT == PRV2_ACT_GELU_TANH) return gelu_tanh(x);
  if (ACT == PRV2_ACT_SIGMOID_GATE) return 1.0f / (1.0f + __expf(-x));     // bi_directional_fusion_model.py:75-77 (gate, then * res below)
  return x;
}

// bulk async copy shared -> global of one staged row segment (row epilogue)
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_f32(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
// TMA tensor store of one staged panel (32 rows x 128 bytes, 128-byte swizzle): out-of-range rows / channels are clipped by the
// tensor map, so ragged edges need no predicates
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t ssrc, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(ssrc), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct Row {              // one output pixel handled by this lane in the coalesced phase
  size_t m, orow;
  int h, w, img;
  bool valid;
};

struct Pre { uint4 a[4]; };   // one chunk's worth of prefetched epilogue operands of this lane (2 rows)

// 8 packed values already in registers (hi plane) [+ lo plane read from memory in the 3-pass precision mode: fp16 pair]
__device__ __forceinline__ void act_unpack8(const uint4& hi, const bf16* lo, size_t i, float (&out)[8]) {
  unpack8(hi, lo != nullptr, out, false);
  if (lo) {
    const uint4 b = *reinterpret_cast<const uint4*>(lo + i);
    unpack8(b, true, out, true);
  }
}

// EPI / ACT are compile-time: each instantiation carries only its own epilogue (small, branch-free SASS)
// FAST selects the row epilogue (plain bf16 output, no residual / ReLU copy / lo plane): lane == TMEM lane == output
// pixel; the lane converts its own row segment, parks it in a padded shared row and ships it with ONE bulk async copy
// (cp.async.bulk shared -> global).  No transposes, no per-chunk global-store instructions, ragged edges by predicate.
template <int EPI, int ACT, int CG, bool FAST>
__global__ void __launch_bounds__(NUM_THREADS, 1) umma_gemm_kernel(const __grid_constant__ KParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stages = p.stages;
  // stage = kb x [A block 16 KB] | kb x [B block]: with a narrow N tile (<= 128) one 64-wide K block is only 256 tensor cycles
  // and the per-stage barrier round trip of the issuing warp became the limit, so two K blocks share a stage there
  const int kb = p.kb;
  const uint32_t blk_bytes = A_STAGE_BYTES + p.b_stage_bytes;           // b_stage_bytes: this CTA's share (half the N tile when CG == 2)
  const uint32_t stage_bytes = p.stage_bytes;
  const uint32_t b_off = p.b_off;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const uint32_t n_workers = CG == 2 ? gridDim.x >> 1 : gridDim.x;      // CTAs (CG 1) or CTA pairs (CG 2) walking the tile list
  const uint32_t worker = CG == 2 ? blockIdx.x >> 1 : blockIdx.x;
  const uint32_t bar_base = smem_base + stages * stage_bytes;
  // barrier map: full[s] | empty[s] | tmem_full[2] | tmem_empty[2] | tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4);
  // epilogue staging behind the barriers (TMA-store mode: 1024-byte aligned for the swizzled panels; bar_base is 1024-aligned there)
  uint8_t* const stage_area = smem_raw + (bar_base - smem_u32(smem_raw)) + (p.tma_store ? 1024 : 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), CG * NUM_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < PRV2_MAX_SRC; ++s) prefetch_tmap(&p.tmA[s]);
    prefetch_tmap(&p.tmB);
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();     // the peer's barriers are initialised before any remote arrive / complete_tx
  else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int tiles_per_img = p.tiles_h * p.tiles_w;

  if (warp == 0) {
    // ================================ TMA producer =========================================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < p.total_tiles; tile += n_workers) {
        const int nt = tile % p.tiles_n, mt = (tile / p.tiles_n) * CG + (int)cta_rank;
        const int img = mt / tiles_per_img, r = mt % tiles_per_img;       // img >= N (odd tile count): the box is out of bounds -> zeros
        const int h0 = (r / p.tiles_w) * p.tile_h, w0 = (r % p.tiles_w) * p.tile_w;
        const int n0 = nt * p.block_n + (int)cta_rank * (p.block_n >> 1) * (CG - 1);
        int kcol = 0;
        for (int s = 0; s < p.n_seg; ++s) {
          const CUtensorMap* map = &p.tmA[p.seg_src[s]];
          const int dh = p.seg_dh[s], dw = p.seg_dw[s];
          const int chunks = p.seg_chunks[s];
          if (p.seg_taps[s] == 3) {
            // vertical tap group: ONE (tile_h + 2)-row halo tile per 64-channel block serves the three dh taps; three weight blocks
            const uint32_t group_bytes = p.halo_bytes + 3 * p.b_stage_bytes;
            for (int c = 0; c < chunks; ++c) {
              mbar_wait(empty_bar(stage), phase ^ 1);
              const uint32_t a_dst = smem_base + stage * stage_bytes;
              if (elect_one()) {
                if ((p.debug & 1) && (phase != 0 || tile != (int)worker)) {
                  if (cta_rank == 0) mbar_arrive(full_bar(stage));
                } else if (CG == 2) {
                  const uint32_t lbar = mapa_shared(full_bar(stage), 0);
                  if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * group_bytes);
                  tma_load_4d_pair(a_dst, map, lbar, c * BK, w0 + dw, h0 + dh - 1, img);
                  for (int r = 0; r < 3; ++r) tma_load_2d_pair(a_dst + b_off + r * p.b_stage_bytes, &p.tmB, lbar, kcol + r * BK, n0);
                } else {
                  mbar_expect_tx(full_bar(stage), group_bytes);
                  tma_load_4d(a_dst, map, full_bar(stage), c * BK, w0 + dw, h0 + dh - 1, img);
                  for (int r = 0; r < 3; ++r) tma_load_2d(a_dst + b_off + r * p.b_stage_bytes, &p.tmB, full_bar(stage), kcol + r * BK, n0);
                }
              }
              __syncwarp();
              kcol += 3 * BK;
              if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            continue;
          }
          for (int c = 0; c < chunks; c += kb) {
            const int nb = min(kb, chunks - c);                 // K blocks of this stage (a segment's odd last block travels alone)
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t a_dst = smem_base + stage * stage_bytes;
            if (elect_one()) {
              if ((p.debug & 1) && (phase != 0 || tile != (int)worker)) {      // diagnostics: operands stay whatever the first ring pass loaded
                if (cta_rank == 0) mbar_arrive(full_bar(stage));
              } else if (CG == 2) {
                // both CTAs' boxes complete on the LEADER's barrier, which expects the bytes of the pair
                const uint32_t lbar = mapa_shared(full_bar(stage), 0);
                if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * nb * blk_bytes);
                for (int j = 0; j < nb; ++j) {
                  tma_load_4d_pair(a_dst + j * A_STAGE_BYTES, map, lbar, (c + j) * BK, w0 + dw, h0 + dh, img);
                  tma_load_2d_pair(a_dst + b_off + j * p.b_stage_bytes, &p.tmB, lbar, kcol + j * BK, n0);
                }
              } else {
                mbar_expect_tx(full_bar(stage), nb * blk_bytes);
                for (int j = 0; j < nb; ++j) {
                  tma_load_4d(a_dst + j * A_STAGE_BYTES, map, full_bar(stage), (c + j) * BK, w0 + dw, h0 + dh, img);
                  tma_load_2d(a_dst + b_off + j * p.b_stage_bytes, &p.tmB, full_bar(stage), kcol + j * BK, n0);
                }
              }
            }
            __syncwarp();
            kcol += nb * BK;
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===========================================
    if (cta_rank == 0) {
      // cute::UMMA::InstrDescriptor: c=F32(1)<<4 | a=BF16(1)<<7 | b=BF16(1)<<10 | N>>3 <<17 | M>>4 <<24, both K-major
      // (CG == 2: M = 256 = 128 rows in each CTA of the pair)
      // c = F32; a / b format: 1 = BF16 (one-pass mode), 0 = F16 (the (hi, lo) pair planes)
      const uint32_t fmt = p.f16 ? 0u : 1u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
      auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t accum) {
        if (CG == 2) tc_mma_bf16_pair(d, a, b, idesc, accum); else tc_mma_bf16(d, a, b, idesc, accum);
      };
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = worker; tile < p.total_tiles; tile += n_workers, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        GTRACE(1, it, 0);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        GTRACE(1, it, 1);
        const uint32_t tmem_d = tmem_base + acc * p.block_n;
        uint32_t accumulate = 0;
        for (int s = 0; s < p.n_seg; ++s) {
          const int chunks = p.seg_chunks[s];
          const int last = p.seg_last[s];
          if (p.seg_taps[s] == 3) {
            for (int c = 0; c < chunks; ++c) {
              const int slices = (c == chunks - 1) ? last : 4;
              mbar_wait(full_bar(stage), phase);
              tc_fence_after();
              const uint32_t a_addr = smem_base + stage * stage_bytes;
              if (elect_one()) {
                for (int r = 0; r < 3; ++r) {
                  // tap dh-1+r reads the halo tile r image rows further down: a 1024-byte-aligned view of the same shared tile
                  const uint64_t adesc = umma_desc_sw128(a_addr + r * p.halo_step), bdesc = umma_desc_sw128(a_addr + b_off + r * p.b_stage_bytes);
                  if (slices == 4) {
                    mma(tmem_d, adesc, bdesc, accumulate);
                    mma(tmem_d, adesc + 2, bdesc + 2, 1);
                    mma(tmem_d, adesc + 4, bdesc + 4, 1);
                    mma(tmem_d, adesc + 6, bdesc + 6, 1);
                  } else {
                    for (int k = 0; k < slices; ++k) mma(tmem_d, adesc + 2 * k, bdesc + 2 * k, k ? 1u : accumulate);
                  }
                  accumulate = 1;
                }
                if (CG == 2) tc_commit_pair(empty_bar(stage)); else tc_commit(empty_bar(stage));
              }
              __syncwarp();
              accumulate = 1;
              if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            continue;
          }
          for (int c = 0; c < chunks; c += kb) {
            const int nb = min(kb, chunks - c);
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_addr = smem_base + stage * stage_bytes;
            if (elect_one()) {
              for (int j = 0; j < nb; ++j) {
                const int slices = (c + j == chunks - 1) ? last : 4;
                const uint64_t adesc = umma_desc_sw128(a_addr + j * A_STAGE_BYTES), bdesc = umma_desc_sw128(a_addr + b_off + j * p.b_stage_bytes);
                // +32 bytes per 16-element K slice inside the 128-byte swizzle row (encoded >>4)
                if (slices == 4) {
                  mma(tmem_d, adesc, bdesc, accumulate);
                  mma(tmem_d, adesc + 2, bdesc + 2, 1);
                  mma(tmem_d, adesc + 4, bdesc + 4, 1);
                  mma(tmem_d, adesc + 6, bdesc + 6, 1);
                } else {
                  for (int k = 0; k < slices; ++k) mma(tmem_d, adesc + 2 * k, bdesc + 2 * k, k ? 1u : accumulate);
                }
                accumulate = 1;
              }
              if (CG == 2) tc_commit_pair(empty_bar(stage)); else tc_commit(empty_bar(stage));
            }
            __syncwarp();
            accumulate = 1;
            if (++stage == stages) { stage = 0; phase ^= 1; }
          }
        }
        if (elect_one()) { if (CG == 2) tc_commit_pair(tfull_bar(acc)); else tc_commit(tfull_bar(acc)); }
        __syncwarp();
        GTRACE(1, it, 2);
      }
    }
  } else if (FAST) {
    // ================================ row epilogue ==========================================
    const int ew = warp - 2;
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
    const int csel = ew >> 2;                     // the two warps of a quadrant alternate 64-column panels
    const int te = threadIdx.x - 64;
    const int Cout = p.Cout, block_n = p.block_n, out_cs = p.out_cs;
    const int tile_w_log2 = p.tile_w_log2, tile_w_mask = p.tile_w - 1, pH = p.H, pW = p.W;
    bf16* const out_hi = p.out_hi;
    // RESID: two 128-byte rows per lane (pitch ROW_PITCH, second buffer at +32*ROW_PITCH); bf16 outputs: one 256-byte row (pitch ROW_PITCH2)
    uint8_t* const srow = stage_area + ew * p.stg_warp_bytes + lane * (EPI == PRV2_EPI_RESID_F32 ? ROW_PITCH : ROW_PITCH2);
    const uint32_t srow_s = smem_u32(srow);
    // TMA-store mode (layers whose epilogue, not their MMAs, paces the tile: per-tile clock64 traces, scripts/gemm_trace.py, show
    // ~2300-3000 clocks per tile for the 32 serialised per-lane bulk copies of a warp, doubled by the second warp of the scheduler):
    // per warp two 4 KB panels (32 rows x 128 B) in the tensor map's 128-byte-swizzled box layout, ONE tensor store per panel
    // issued by one lane.  (For the MMA-bound ViT linears the tensor stores compete with the operand loads for the TMA unit:
    // measured 137 -> 147 us at qkv, so those keep the per-lane copies.)
    const bool tma = p.tma_store != 0;
    uint8_t* const wstage = stage_area + ew * p.stg_warp_bytes;
    const uint32_t wstage_s = smem_u32(wstage);
    uint8_t* const lrow = wstage + lane * 128;
    const uint32_t lsw = (uint32_t)(lane & 7);
    const int bq_w = (quad * 32) & tile_w_mask, bq_h = (quad * 32) >> tile_w_log2;      // this warp's box origin inside the tile
    float* const s_par = reinterpret_cast<float*>(stage_area + NUM_EPI_WARPS * p.stg_warp_bytes);
    float* const s_ln = reinterpret_cast<float*>(stage_area + NUM_EPI_WARPS * p.stg_warp_bytes + PAR_BYTES);
    int n_sub_done = 0;
    constexpr bool is_ln = EPI == PRV2_EPI_LN_GELU;
    // the two warps of a quadrant own contiguous halves of the tile's 16-column chunks: ONE bulk copy per lane and tile
    const int n16 = block_n >> 4, q_half = (n16 + 1) >> 1;
    const int q_start = csel ? q_half : 0, q_cnt = csel ? n16 - q_half : q_half;
    const uint32_t tempty_leader0 = CG == 2 ? mapa_shared(tempty_bar(0), 0) : tempty_bar(0);
    int it = 0;
    for (int tile = worker; tile < p.total_tiles; tile += n_workers, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int nt = tile % p.tiles_n, mt = (tile / p.tiles_n) * CG + (int)cta_rank;
      const int img = mt / tiles_per_img, r = mt % tiles_per_img;
      const int h0 = (r / p.tiles_w) * p.tile_h, w0 = (r % p.tiles_w) * p.tile_w;
      const int n0 = nt * block_n;
      const int tr = quad * 32 + lane;
      const int h = h0 + (tr >> tile_w_log2), w = w0 + (tr & tile_w_mask);
      const bool valid = (h < pH) && (w < pW) && (img < p.N);
      bf16* const grow = out_hi + (((size_t)img * pH + h) * pW + w) * out_cs + n0;
      // per-tile parameters (bias / gamma / beta of this N tile) -> shared memory.  With a single N tile they are the same for every
      // tile: loaded once (first iteration, both parity buffers), which also frees the eight epilogue warps from meeting at a
      // barrier on every tile (per-tile traces: ~150-500 clocks of global-load latency + lockstep per tile)
      float* const par = s_par + (p.tiles_n == 1 ? 0 : acc * 768);
      if (ew == 0) GTRACE(0, it, 0);
      if (p.tiles_n != 1 || it == 0) {
        if (te < block_n) {
          const int n = min(n0 + te, Cout - 1);
          par[te] = p.bias ? __ldg(p.bias + n) : 0.f;
          if (is_ln) { par[256 + te] = __ldg(p.gamma + n); par[512 + te] = __ldg(p.beta + n); }
          if (EPI == PRV2_EPI_RESID_F32) par[256 + te] = __ldg(p.gamma + n);
        }
        asm volatile("bar.sync 5, 256;" ::: "memory");        // parameters visible to all epilogue warps
      }
      if (ew == 0) GTRACE(0, it, 1);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (ew == 0) GTRACE(0, it, 2);
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * block_n;
      float mean = 0.f, rstd = 1.f;
      if (is_ln) {
        // channels-first LayerNorm statistics of this lane's pixel (convs.py:24-27): each warp of the quadrant reduces its own
        // panels in ONE pass over tensor memory -- sums of (x - shift) and (x - shift)^2 with shift = the first channel of the
        // warp's range, which keeps the cancellation in s2 - s1^2 / n at the (shift - mean)^2 / var level (< 1e-5 relative on the
        // variance for a shift within 10 sigma) -- the load of chunk q + 1 in flight while chunk q is reduced; the halves merge
        // with Chan's parallel update.  (Round 1 read the accumulator three times with a blocking wait per chunk; at K = 1152 +
        // depth taps the epilogue, not the MMAs, paced FusionUnet's encoder_layers_2[0]: 701 TFLOP/s in-step.)
        float s1 = 0.f, s2 = 0.f, shift = 0.f;
        int cnt = 0;
        uint32_t ra[16], rb[16];
        if (q_cnt > 0) { tc_ld16_issue(taddr + q_start * 16, ra); tc_ld_wait(); }
        for (int q = 0; q < q_cnt; ++q) {
          if (q + 1 < q_cnt) tc_ld16_issue(taddr + (q_start + q + 1) * 16, rb);
          const int c0 = (q_start + q) * 16;
          if (q == 0) shift = fmaf(__uint_as_float(ra[0]), p.acc_scale, par[c0]);
          if (n0 + c0 + 16 <= Cout) {                          // whole chunk valid: four independent accumulation chains, no predicates
            float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float4 b4 = *reinterpret_cast<const float4*>(par + c0 + g * 4);
              const float d0 = fmaf(__uint_as_float(ra[g * 4 + 0]), p.acc_scale, b4.x) - shift, d1 = fmaf(__uint_as_float(ra[g * 4 + 1]), p.acc_scale, b4.y) - shift;
              const float d2 = fmaf(__uint_as_float(ra[g * 4 + 2]), p.acc_scale, b4.z) - shift, d3 = fmaf(__uint_as_float(ra[g * 4 + 3]), p.acc_scale, b4.w) - shift;
              a1[0] += d0; a1[1] += d1; a1[2] += d2; a1[3] += d3;
              a2[0] = fmaf(d0, d0, a2[0]); a2[1] = fmaf(d1, d1, a2[1]); a2[2] = fmaf(d2, d2, a2[2]); a2[3] = fmaf(d3, d3, a2[3]);
            }
            s1 += (a1[0] + a1[1]) + (a1[2] + a1[3]);
            s2 += (a2[0] + a2[1]) + (a2[2] + a2[3]);
            cnt += 16;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c0 + j < Cout) { const float d = fmaf(__uint_as_float(ra[j]), p.acc_scale, par[c0 + j]) - shift; s1 += d; s2 = fmaf(d, d, s2); ++cnt; }
          }
          if (q + 1 < q_cnt) {
            tc_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) ra[j] = rb[j];
          }
        }
        const float inv_cnt = cnt ? 1.0f / (float)cnt : 0.f;
        const float mean_a = cnt ? fmaf(s1, inv_cnt, shift) : 0.f;
        const float m2 = fmaxf(s2 - s1 * s1 * inv_cnt, 0.f);
        float* const mine = s_ln + (ew * 2 + acc) * 64;
        const float* const theirs = s_ln + ((ew ^ 4) * 2 + acc) * 64;
        mine[lane] = mean_a;
        mine[32 + lane] = m2;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        const float mean_b = theirs[lane], m2_b = theirs[32 + lane];
        const float na = (float)cnt, nb = (float)(Cout - cnt), nn = (float)Cout;
        const float delta = mean_b - mean_a;
        mean = mean_a + delta * (nb / nn);
        rstd = 1.0f / sqrtf((m2 + m2_b + delta * delta * (na * nb / nn)) / nn + p.eps);
      }
      if (EPI == PRV2_EPI_RESID_F32) {
        // x[row, n] += gamma[n] * (acc + bias[n])  (block.py:105-106, layer_scale.py:27): the lane stages 32 fp32 columns
        // of its row and hands them to the L2 as ONE bulk reduce-add -- the SM never reads the residual stream, so the
        // read-modify-write latency that bound this epilogue is gone.  Two staged rows per lane alternate.
        float* const grow32 = p.out_f32 + (((size_t)img * pH + h) * pW + w) * p.out_f32_ld + n0;
        const int n_sub = (block_n + 31) >> 5;
        for (int sp = csel; sp < n_sub; sp += 2, ++n_sub_done) {
          const int c0 = sp * 32;
          const int cols = min(32, block_n - c0);
          const int buf = n_sub_done & 1;
          bulk_wait_read1();                                  // the copy issued two steps ago has left this buffer
          uint32_t ra[16], rb[16];
          tc_ld16_issue(taddr + c0, ra);
          if (cols > 16) tc_ld16_issue(taddr + c0 + 16, rb);
          tc_ld_wait();
          uint8_t* const dst = srow + buf * (32 * ROW_PITCH);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (g * 4 >= cols) break;
            const float4 b4 = *reinterpret_cast<const float4*>(par + c0 + g * 4), g4 = *reinterpret_cast<const float4*>(par + 256 + c0 + g * 4);
            const uint32_t* rsrc = g < 4 ? &ra[g * 4] : &rb[(g - 4) * 4];
            float4 o;
            const float sc = p.acc_scale;
            o.x = g4.x * fmaf(__uint_as_float(rsrc[0]), sc, b4.x); o.y = g4.y * fmaf(__uint_as_float(rsrc[1]), sc, b4.y);
            o.z = g4.z * fmaf(__uint_as_float(rsrc[2]), sc, b4.z); o.w = g4.w * fmaf(__uint_as_float(rsrc[3]), sc, b4.w);
            *reinterpret_cast<float4*>(dst + g * 16) = o;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          const int nvalid = min(cols, Cout - (n0 + c0));
          if (valid && nvalid > 0 && !(p.debug & 4)) bulk_reduce_add_f32(grow32 + c0, srow_s + buf * (32 * ROW_PITCH), (uint32_t)nvalid * 4u);
          bulk_commit();
        }
      } else
      if (q_cnt > 0) {
        const int c_first = q_start * 16;
        if (tma) { if (lane == 0) bulk_wait_read0(); __syncwarp(); }   // the stores of the previous tile have read the panels
        else bulk_wait_read0();                               // this lane's copy of the previous tile has left its staging row
        if (ew == 0) GTRACE(0, it, 3);
        uint32_t rr[16], rn[16];
        tc_ld16_issue(taddr + c_first, rr);
        tc_ld_wait();
        for (int q = 0; q < q_cnt; ++q) {
          if (q + 1 < q_cnt) tc_ld16_issue(taddr + c_first + (q + 1) * 16, rn);
          const int nl = c_first + q * 16;
          float t[16];
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 b4 = *reinterpret_cast<const float4*>(par + nl + g * 4);
            t[g * 4 + 0] = __uint_as_float(rr[g * 4 + 0]); t[g * 4 + 1] = __uint_as_float(rr[g * 4 + 1]);
            t[g * 4 + 2] = __uint_as_float(rr[g * 4 + 2]); t[g * 4 + 3] = __uint_as_float(rr[g * 4 + 3]);
            if (is_ln) {
              const float4 g4 = *reinterpret_cast<const float4*>(par + 256 + nl + g * 4), e4 = *reinterpret_cast<const float4*>(par + 512 + nl + g * 4);
              t[g * 4 + 0] = g4.x * ((t[g * 4 + 0] + b4.x - mean) * rstd) + e4.x; t[g * 4 + 1] = g4.y * ((t[g * 4 + 1] + b4.y - mean) * rstd) + e4.y;
              t[g * 4 + 2] = g4.z * ((t[g * 4 + 2] + b4.z - mean) * rstd) + e4.z; t[g * 4 + 3] = g4.w * ((t[g * 4 + 3] + b4.w - mean) * rstd) + e4.w;
            } else {
              t[g * 4 + 0] += b4.x; t[g * 4 + 1] += b4.y; t[g * 4 + 2] += b4.z; t[g * 4 + 3] += b4.w;
            }
          }
          __nv_bfloat162 h2[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) h2[j] = __floats2bfloat162_rn(act_fn<ACT>(t[2 * j]), act_fn<ACT>(t[2 * j + 1]));
          if (tma) {                                          // chunk q = 32 bytes of panel q / 4, swizzled 16-byte pieces
            uint8_t* const pb = lrow + (q >> 2) * 4096;
            const uint32_t j0 = (uint32_t)(q & 3) * 2;
            *reinterpret_cast<uint4*>(pb + ((j0 ^ lsw) << 4)) = *reinterpret_cast<const uint4*>(&h2[0]);
            *reinterpret_cast<uint4*>(pb + (((j0 + 1) ^ lsw) << 4)) = *reinterpret_cast<const uint4*>(&h2[4]);
          } else {
            *reinterpret_cast<uint4*>(srow + q * 32) = *reinterpret_cast<const uint4*>(&h2[0]);
            *reinterpret_cast<uint4*>(srow + q * 32 + 16) = *reinterpret_cast<const uint4*>(&h2[4]);
          }
          if (q + 1 < q_cnt) {
            tc_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) rr[j] = rn[j];
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // staged row -> visible to the bulk copy engine
        const int nvalid = min(q_cnt * 16, Cout - (n0 + c_first));
        if (ew == 0) GTRACE(0, it, 4);
        if (tma) {
          __syncwarp();                                        // all 32 rows of the panels are staged
          if (lane == 0) {
            for (int pn = 0; pn * 64 < q_cnt * 16; ++pn)
              if (n0 + c_first + pn * 64 < Cout && !(p.debug & 4))
                tma_store_4d(&p.tmOut, wstage_s + pn * 4096, n0 + c_first + pn * 64, w0 + bq_w, h0 + bq_h, img);
            bulk_commit();
          }
        } else {
          if (valid && nvalid > 0 && !(p.debug & 4)) bulk_store(grow + c_first, srow_s, (uint32_t)nvalid * 2u);
          bulk_commit();
        }
      }
      if (ew == 0) GTRACE(0, it, 5);
      tc_fence_before();
      __syncwarp();                                          // every lane's TMEM reads of this accumulator have completed
      if (ew == 0) GTRACE(0, it, 6);
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty_leader0 + 8u * acc);
        else mbar_arrive(tempty_bar(acc));
      }
      if (ew == 0) GTRACE(0, it, 7);
    }
    bulk_wait_all();
  } else {
    // ================================ epilogue =============================================
    // Phase A: the lane that owns TMEM lane (= output pixel) r pulls 16 fp32 columns and parks them in a
    // padded shared tile.  Phase B: lanes re-read the tile so that a lane owns 8 CONSECUTIVE channels of
    // rows (lane&15) and (lane&15)+16 -> every global access is a 16-byte vector and a warp instruction
    // covers whole 32-byte sectors.  Two warps share a lane quadrant and alternate 16-column chunks.
    const int ew = warp - 2;
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
    const int Cout = p.Cout, block_n = p.block_n, out_cs = p.out_cs, relu_cs = p.relu_cs, res_cs = p.res_cs, res2_cs = p.res2_cs;
    const int out_f32_ld = p.out_f32_ld, tile_w_log2 = p.tile_w_log2, tile_w_mask = p.tile_w - 1, pH = p.H, pW = p.W;
    bf16* const out_hi = p.out_hi; bf16* const out_lo = p.out_lo;
    bf16* const relu_hi = p.relu_hi; bf16* const relu_lo = p.relu_lo;
    const bf16* const res_hi = p.res_hi; const bf16* const res_lo = p.res_lo;
    const bf16* const res2_hi = p.res2_hi; const bf16* const res2_lo = p.res2_lo;
    float* const out_f32 = p.out_f32;
    const int csel = ew >> 2;                     // which of the two warps of this quadrant
    const int te = threadIdx.x - 64;              // 0..255 over the epilogue warps
    const int n_chunks = block_n >> 4;
    float* const stg = reinterpret_cast<float*>(stage_area + ew * p.stg_warp_bytes);
    float* const stg_peer = reinterpret_cast<float*>(stage_area + (ew ^ 4) * p.stg_warp_bytes);
    float* const s_par = reinterpret_cast<float*>(stage_area + NUM_EPI_WARPS * p.stg_warp_bytes);
    const int rB = lane & 15, cB = (lane >> 4) * 8;
    constexpr bool is_ln = EPI == PRV2_EPI_LN_GELU;
    const uint32_t tempty_leader0 = CG == 2 ? mapa_shared(tempty_bar(0), 0) : tempty_bar(0);
    int it = 0;
    for (int tile = worker; tile < p.total_tiles; tile += n_workers, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int nt = tile % p.tiles_n, mt = (tile / p.tiles_n) * CG + (int)cta_rank;
      const int img = mt / tiles_per_img, r = mt % tiles_per_img;
      const int h0 = (r / p.tiles_w) * p.tile_h, w0 = (r % p.tiles_w) * p.tile_w;
      const int n0 = nt * block_n;
      Row rows[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int tr = quad * 32 + rB + 16 * i;
        Row& q = rows[i];
        q.img = img;
        q.h = h0 + (tr >> tile_w_log2);
        q.w = w0 + (tr & tile_w_mask);
        q.valid = (q.h < pH) && (q.w < pW) && (img < p.N);
        q.m = ((size_t)img * pH + q.h) * pW + q.w;
        q.orow = q.m;
        if (p.row_map_period > 0) q.orow = q.m + (q.m / p.row_map_period) * p.row_map_extra + p.row_map_offset;
      }
      // (1) per-tile parameters -> shared memory (one coalesced load instead of 8..24 scalar loads per chunk)
      float* const par = s_par + acc * 768;
      if (te < block_n) {
        const int n = min(n0 + te, Cout - 1);
        int nb = n;
        if (EPI == PRV2_EPI_SHUFFLE) nb = n % (Cout / (p.shuffle_k * p.shuffle_k));
        par[te] = p.bias ? __ldg(p.bias + nb) : 0.f;
        par[256 + te] = p.gamma ? __ldg(p.gamma + n) : 0.f;
        par[512 + te] = (p.beta && is_ln) ? __ldg(p.beta + n) : 0.f;
      }
      // (2) operands the epilogue READS from global memory (fp32 residual stream / residual activations) do not
      //     depend on the accumulator: fetch them into registers three chunks ahead, starting NOW, while the MMAs
      //     of this tile are still running.  (A load issued only when its chunk is processed costs a full
      //     DRAM/L2 latency per row and made the read-modify-write epilogues latency-bound.)
      constexpr bool kPre = (EPI == PRV2_EPI_RESID_F32) || (EPI == PRV2_EPI_STORE);
      const bool x_vec = (out_f32_ld & 3) == 0 && (Cout & 7) == 0;
      auto prefetch = [&](int cc, Pre& pre) {
        if (!kPre || cc >= n_chunks) return;
        const int n = n0 + cc * 16 + cB;
        if (n >= Cout) return;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const Row& q = rows[i];
          if (!q.valid) continue;
          if (EPI == PRV2_EPI_RESID_F32) {
            if (x_vec) {
              const uint4* x = reinterpret_cast<const uint4*>(out_f32 + q.orow * out_f32_ld + n);
              pre.a[2 * i] = x[0]; pre.a[2 * i + 1] = x[1];
            }
          } else {
            if (res_hi) pre.a[i] = *reinterpret_cast<const uint4*>(res_hi + q.m * res_cs + n);
            if (res2_hi) pre.a[2 + i] = *reinterpret_cast<const uint4*>(res2_hi + q.m * res2_cs + n);
          }
        }
      };
      Pre pre0, pre1, pre2;
      prefetch(csel, pre0);
      if (EPI == PRV2_EPI_RESID_F32) { prefetch(csel + 2, pre1); prefetch(csel + 4, pre2); }
      asm volatile("bar.sync 5, 256;" ::: "memory");          // parameters visible to all epilogue warps
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * block_n;
      float v[16];

      if (p.debug & 2) {
        tc_ld16(taddr + csel * 16, v, 1.0f);
        if (v[0] == 123.456f && out_f32) out_f32[0] = v[1];             // keep the load alive
      } else if (EPI == PRV2_EPI_HEAD) {
        if (csel == 0) {                           // one float per pixel: lane-per-row stores are already coalesced
          const int tr = quad * 32 + lane;
          const int h = h0 + (tr >> tile_w_log2), w = w0 + (tr & tile_w_mask);
          float dot = 0.f;
          for (int c = 0; c < n_chunks; ++c) {
            tc_ld16(taddr + c * 16, v, p.acc_scale);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c * 16 + j < Cout) dot += fmaxf(v[j] + par[c * 16 + j], 0.f) * par[256 + c * 16 + j];
          }
          if (h < pH && w < pW && img < p.N) out_f32[((size_t)img * pH + h) * pW + w] = p.head_scale / (1.0f + expf(-(dot + __ldg(p.beta))));
        }
      } else {
        float* const s_fin = stg + 32 * STG_LD + 128;          // [mean(32) | rstd(32)]
        if (is_ln) {
          // channels-first LayerNorm statistics of this lane's pixel (convs.py:24-27).  The two warps of the
          // quadrant each reduce their own chunks two-pass (mean, then centred squares) and merge with
          // Chan's parallel update; partials are double buffered by accumulator parity.
          float* const s_mine = stg + 32 * STG_LD + acc * 64;
          float* const s_theirs = stg_peer + 32 * STG_LD + acc * 64;
          float sum = 0.f;
          int cnt = 0;
          for (int c = csel; c < n_chunks; c += 2) {
            tc_ld16(taddr + c * 16, v, p.acc_scale);
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n0 + c * 16 + j < Cout) { sum += v[j] + par[c * 16 + j]; ++cnt; }
          }
          const float mean_a = cnt ? sum / (float)cnt : 0.f;
          float m2 = 0.f;
          for (int c = csel; c < n_chunks; c += 2) {
            tc_ld16(taddr + c * 16, v, p.acc_scale);
#pragma unroll
            for (int j = 0; j < 16; ++j) if (n0 + c * 16 + j < Cout) { const float d = v[j] + par[c * 16 + j] - mean_a; m2 += d * d; }
          }
          s_mine[lane] = mean_a;
          s_mine[32 + lane] = m2;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
          const float mean_b = s_theirs[lane], m2_b = s_theirs[32 + lane];
          const float na = (float)cnt, nb = (float)(Cout - cnt), nn = (float)Cout;
          const float delta = mean_b - mean_a;
          const float mean = mean_a + delta * (nb / nn);
          const float var = (m2 + m2_b + delta * delta * (na * nb / nn)) / nn;
          s_fin[lane] = mean;
          s_fin[32 + lane] = 1.0f / sqrtf(var + p.eps);
        }
        // (3) chunk loop, software pipelined: the TMEM load of chunk c+2 is in flight while chunk c goes
        //     through its coalesced phase.
        if (csel < n_chunks) tc_ld16(taddr + csel * 16, v, p.acc_scale);
        auto chunk = [&](const int c, const Pre& pre) {
          float4* dst = reinterpret_cast<float4*>(stg + lane * STG_LD);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
          dst[2] = make_float4(v[8], v[9], v[10], v[11]);
          dst[3] = make_float4(v[12], v[13], v[14], v[15]);
          __syncwarp();                                      // tile visible; warp converged for the .aligned load below
          const bool more = c + 2 < n_chunks;
          uint32_t rr[16];
          if (more) {
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]), "=r"(rr[8]),
                  "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
                : "r"(taddr + (c + 2) * 16)
                : "memory");
          }
          const int nl = c * 16 + cB;                        // first of this lane's 8 channels, tile-local
          const int n = n0 + nl;
          if (n < Cout && !(p.debug & 8)) {
            const float4 b0 = *reinterpret_cast<const float4*>(par + nl), b1 = *reinterpret_cast<const float4*>(par + nl + 4);
            const float4 g0 = *reinterpret_cast<const float4*>(par + 256 + nl), g1 = *reinterpret_cast<const float4*>(par + 256 + nl + 4);
            const float bias8[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            const float g8[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            int sh_co = 0, sh_ky = 0, sh_kx = 0;
            if (EPI == PRV2_EPI_SHUFFLE) {
              // n = (ky*k + kx)*C + co ; Cout counts k*k*C columns
              const int k = p.shuffle_k, co_n = Cout / (k * k), tap = n / co_n;
              sh_co = n % co_n; sh_ky = tap / k; sh_kx = tap % k;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const Row& q = rows[i];
              if (!q.valid) continue;
              const int rl = rB + 16 * i;
              const float4 lo4 = *reinterpret_cast<const float4*>(stg + rl * STG_LD + cB);
              const float4 hi4 = *reinterpret_cast<const float4*>(stg + rl * STG_LD + cB + 4);
              float t[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
              if (is_ln) {
                const float4 e0 = *reinterpret_cast<const float4*>(par + 512 + nl), e1 = *reinterpret_cast<const float4*>(par + 512 + nl + 4);
                const float b8[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                const float mean = s_fin[rl], rstd = s_fin[32 + rl];
#pragma unroll
                for (int j = 0; j < 8; ++j) t[j] = act_fn<ACT>(g8[j] * ((t[j] + bias8[j] - mean) * rstd) + b8[j]);
                act_store8(out_hi, out_lo, q.orow * out_cs + n, t);
                continue;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) t[j] += bias8[j];
              if (EPI == PRV2_EPI_SHUFFLE) {
                if (ACT != PRV2_ACT_NONE) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) t[j] = act_fn<ACT>(t[j]);
                }
                const int k = p.shuffle_k;
                const size_t opix = ((size_t)q.img * (pH * k) + (q.h * k + sh_ky)) * (pW * k) + (q.w * k + sh_kx);
                act_store8(out_hi, out_lo, opix * out_cs + sh_co, t);
                continue;
              }
              if (EPI == PRV2_EPI_RESID_F32) {
                float* x = out_f32 + q.orow * out_f32_ld + n;
                if (x_vec) {
                  float4 x0 = *reinterpret_cast<const float4*>(&pre.a[2 * i]), x1 = *reinterpret_cast<const float4*>(&pre.a[2 * i + 1]);
                  x0.x += g8[0] * t[0]; x0.y += g8[1] * t[1]; x0.z += g8[2] * t[2]; x0.w += g8[3] * t[3];
                  x1.x += g8[4] * t[4]; x1.y += g8[5] * t[5]; x1.z += g8[6] * t[6]; x1.w += g8[7] * t[7];
                  *reinterpret_cast<float4*>(x) = x0;
                  *reinterpret_cast<float4*>(x + 4) = x1;
                } else {
                  for (int j = 0; j < 8 && n + j < Cout; ++j) x[j] += g8[j] * t[j];
                }
                continue;
              }
              if (EPI == PRV2_EPI_F32) {
                float* x = out_f32 + q.orow * out_f32_ld + n;
                if (n + 8 <= Cout && (out_f32_ld & 3) == 0) {
                  *reinterpret_cast<float4*>(x) = make_float4(t[0], t[1], t[2], t[3]);
                  *reinterpret_cast<float4*>(x + 4) = make_float4(t[4], t[5], t[6], t[7]);
                } else {
                  for (int j = 0; j < 8 && n + j < Cout; ++j) x[j] = t[j];
                }
                continue;
              }
              if (ACT != PRV2_ACT_NONE) {
#pragma unroll
                for (int j = 0; j < 8; ++j) t[j] = act_fn<ACT>(t[j]);
              }
              if (res_hi) {
                float r8[8];
                act_unpack8(pre.a[i], res_lo, q.m * res_cs + n, r8);
#pragma unroll
                for (int j = 0; j < 8; ++j) t[j] = (ACT == PRV2_ACT_SIGMOID_GATE) ? t[j] * r8[j] : t[j] + r8[j];
              }
              if (res2_hi) {
                float r8[8];
                act_unpack8(pre.a[2 + i], res2_lo, q.m * res2_cs + n, r8);
#pragma unroll
                for (int j = 0; j < 8; ++j) t[j] += r8[j];
              }
              if (out_hi && !(p.debug & 4)) act_store8(out_hi, out_lo, q.orow * out_cs + n, t);
              if (relu_hi) {
#pragma unroll
                for (int j = 0; j < 8; ++j) t[j] = fmaxf(t[j], 0.f);
                act_store8(relu_hi, relu_lo, q.orow * relu_cs + n, t);
              }
            }
          }
          __syncwarp();                                      // coalesced phase done (tile reusable), lanes reconverged
          if (more) {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]) * p.acc_scale;
          }
        };
        if (EPI == PRV2_EPI_RESID_F32) {
          // three-deep register ring: the global reads of chunk c+6 are issued right after chunk c retires
          for (int c = csel; c < n_chunks; c += 6) {
            chunk(c, pre0); prefetch(c + 6, pre0);
            if (c + 2 < n_chunks) { chunk(c + 2, pre1); prefetch(c + 8, pre1); }
            if (c + 4 < n_chunks) { chunk(c + 4, pre2); prefetch(c + 10, pre2); }
          }
        } else if (EPI == PRV2_EPI_STORE) {
          // residual activations (16 B per row): issue the reads of the NEXT chunk before working on this one.
          // (A code-doubling two-buffer ring measurably slowed the residual-free GELU epilogue of fc1.)
          for (int c = csel; c < n_chunks; c += 2) {
            if (res_hi) { pre1 = pre0; prefetch(c + 2, pre0); }
            chunk(c, pre1);
          }
        } else {
          for (int c = csel; c < n_chunks; c += 2) chunk(c, pre0);
        }
      }
      tc_fence_before();
      __syncwarp();                                          // every lane's TMEM reads of this accumulator have completed
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty_leader0 + 8u * acc);
        else mbar_arrive(tempty_bar(acc));
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();     // no CTA of the pair exits (or frees TMEM) while the other may still signal it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
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

template <int EPI, int ACT, int CG, bool FAST>
cudaError_t launch(const cudaLaunchConfig_t& cfg, const KParams& p) {
  static bool attr_done = false;                   // per instantiation (one device per process)
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm_kernel<EPI, ACT, CG, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return cudaLaunchKernelEx(&cfg, umma_gemm_kernel<EPI, ACT, CG, FAST>, p);
}

}  // namespace

#ifdef PRV2_GEMM_TRACE_BUILD
extern "C" int prv2_debug_gemm_trace(unsigned long long* out /*host, 3*64*8*/) {
  PRV2_CUDA(cudaDeviceSynchronize());
  PRV2_CUDA(cudaMemcpyFromSymbol(out, g_gemm_trace, sizeof(unsigned long long) * 3 * 64 * 8));
  return PRV2_OK;
}
#endif

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
  // sources read through vertical tap groups are fetched as tiles with a one-row halo above and below
  bool src_halo[PRV2_MAX_SRC] = {false}, src_single[PRV2_MAX_SRC] = {false}, any_halo = false;
  for (int s = 0; s < d->n_seg; ++s) {
    const prv2_seg& sg = d->seg[s];
    PRV2_CHECK_ARG(sg.src >= 0 && sg.src < d->n_src, "prv2_umma_gemm: segment %d references source %d", s, sg.src);
    PRV2_CHECK_ARG(sg.taps_h == 0 || sg.taps_h == 1 || sg.taps_h == 3, "prv2_umma_gemm: segment %d: taps_h must be 1 or 3", s);
    if (sg.taps_h == 3) { src_halo[sg.src] = true; any_halo = true; } else src_single[sg.src] = true;
  }
  for (int s = 0; s < d->n_src; ++s)
    PRV2_CHECK_ARG(!(src_halo[s] && src_single[s]), "prv2_umma_gemm: source %d is used by both tap groups and single taps", s);
  for (int s = 0; s < d->n_src; ++s) {
    const prv2_src& src = d->src[s];
    PRV2_CHECK_ARG(src.ptr && src.C > 0 && src.cs >= src.C && src.cs % 8 == 0 && ((uintptr_t)src.ptr & 15) == 0,
                   "prv2_umma_gemm: source %d needs C>0, pitch multiple of 8, 16-byte aligned", s);
    cuuint64_t dims[4] = {(cuuint64_t)src.C, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)src.cs * 2, (cuuint64_t)d->W * src.cs * 2, (cuuint64_t)d->H * d->W * src.cs * 2};
    cuuint32_t box[4] = {BK, (cuuint32_t)d->tile_w, (cuuint32_t)(d->tile_h + (src_halo[s] ? 2 : 0)), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.tmA[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)src.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("prv2_umma_gemm: cuTensorMapEncodeTiled(A%d) failed (%d) C=%d cs=%d N=%d H=%d W=%d", s, (int)r, src.C, src.cs, d->N, d->H, d->W); return PRV2_ECUDA; }
  }
  for (int s = d->n_src; s < PRV2_MAX_SRC; ++s) p.tmA[s] = p.tmA[0];
  // CTA pairs (cta_group::2) whenever there are two M tiles: each CTA of a pair stages its own
  // A tile and half of the weight tile.  PRV2_GEMM_CG=1|2 forces the choice (diagnostics).
  const int tiles_w = cdiv(d->W, d->tile_w), tiles_h = cdiv(d->H, d->tile_h);
  const long long m_tiles = (long long)d->N * tiles_h * tiles_w;
  const int tiles_n = d->Cout_pad / d->block_n;
  static const char* cg_env = getenv("PRV2_GEMM_CG");
  int cg = m_tiles >= 2 ? 2 : 1;            // a pair runs two M tiles side by side, so small problems lose no parallelism
  if (cg_env && (cg_env[0] == '1' || cg_env[0] == '2')) cg = cg_env[0] - '0';
  if (m_tiles < 2 || d->block_n % 16 != 0) cg = 1;
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)d->Ktot * 2};
    cuuint32_t box[2] = {BK, (cuuint32_t)(d->block_n / cg)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)d->weight, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("prv2_umma_gemm: cuTensorMapEncodeTiled(B) failed (%d) Ktot=%d Cout_pad=%d", (int)r, d->Ktot, d->Cout_pad); return PRV2_ECUDA; }
  }
  int ktot = 0;
  for (int s = 0; s < d->n_seg; ++s) {
    const prv2_seg& sg = d->seg[s];
    const int C = d->src[sg.src].C;
    const int chunks = (C + BK - 1) / BK;
    PRV2_CHECK_ARG(chunks <= 255, "prv2_umma_gemm: segment too long");
    p.seg_src[s] = sg.src; p.seg_dh[s] = sg.dh; p.seg_dw[s] = sg.dw;
    p.seg_chunks[s] = (uint8_t)chunks;
    p.seg_taps[s] = sg.taps_h == 3 ? 3 : 1;
    const int tail = C - (chunks - 1) * BK;
    p.seg_last[s] = (uint8_t)((tail + 15) / 16);
    ktot += chunks * BK * p.seg_taps[s];
  }
  PRV2_CHECK_ARG(ktot == d->Ktot, "prv2_umma_gemm: Ktot %d does not match the segment table (%d)", d->Ktot, ktot);
  p.n_seg = d->n_seg;
  p.N = d->N; p.H = d->H; p.W = d->W; p.Cout = d->Cout;
  p.tile_w = d->tile_w; p.tile_h = d->tile_h;
  PRV2_CHECK_ARG((d->tile_w & (d->tile_w - 1)) == 0, "prv2_umma_gemm: tile_w must be a power of two");
  p.tile_w_log2 = 0;
  while ((1 << p.tile_w_log2) < d->tile_w) ++p.tile_w_log2;
  p.tiles_w = tiles_w; p.tiles_h = tiles_h;
  p.tiles_n = tiles_n;
  const long long total = ((m_tiles + cg - 1) / cg) * p.tiles_n;         // tiles (CG 1) or tile pairs (CG 2)
  PRV2_CHECK_ARG(total < (1LL << 31), "prv2_umma_gemm: too many tiles");
  p.total_tiles = (int)total;
  p.block_n = d->block_n;
  p.b_stage_bytes = (d->block_n / cg) * BK * 2;
  static const char* kb_env = getenv("PRV2_GEMM_KB");             // diagnostics: force the K blocks per stage
  p.kb = d->block_n <= 208 ? 2 : 1;
  if (kb_env && (kb_env[0] == '1' || kb_env[0] == '2')) p.kb = kb_env[0] - '0';
  if (any_halo) p.kb = 1;
  p.halo_step = d->tile_w * BK * 2;                                         // one image row of the tile (a multiple of 1024 B: tile_w >= 16)
  p.halo_bytes = any_halo ? (d->tile_h + 2) * p.halo_step : 0;
  PRV2_CHECK_ARG(!any_halo || d->tile_w >= 8, "prv2_umma_gemm: tap groups need tile_w >= 8");
  p.b_off = any_halo ? p.halo_bytes : p.kb * A_STAGE_BYTES;
  p.stage_bytes = p.b_off + (any_halo ? 3 : p.kb) * p.b_stage_bytes;
  p.stages = SMEM_BUDGET / p.stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  static const char* dbg_env = getenv("PRV2_GEMM_DEBUG");
  p.debug = dbg_env ? atoi(dbg_env) : 0;
  p.stg_warp_bytes = STG_BYTES_PER_WARP;
  static const char* st_env = getenv("PRV2_GEMM_STAGES");          // diagnostics: cap the pipeline depth
  if (st_env && atoi(st_env) >= 2 && atoi(st_env) < p.stages) p.stages = atoi(st_env);
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
  p.acc_scale = d->acc_scale == 0.f ? 1.f : d->acc_scale;
  p.f16 = d->f16;
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
  }
  // row epilogue (bulk-copy stores) whenever the layer writes one plain bf16 tensor
  static const char* fast_env = getenv("PRV2_GEMM_FAST");
  bool fast = !(fast_env && fast_env[0] == '0') && d->out_hi && !d->out_lo && d->row_map_period == 0 && d->out_cs % 8 == 0 && d->Cout % 8 == 0 &&
              ((d->epi == PRV2_EPI_STORE && !d->relu_hi && !d->res_hi && !d->res2_hi) || d->epi == PRV2_EPI_LN_GELU);
  static const char* red_env = getenv("PRV2_GEMM_REDUCE");          // diagnostics: 0 keeps the register read-modify-write epilogue
  const bool fast_resid = d->epi == PRV2_EPI_RESID_F32 && !(fast_env && fast_env[0] == '0') && !(red_env && red_env[0] == '0') &&
                          d->row_map_period == 0 && d->out_f32_ld % 4 == 0 && d->Cout % 4 == 0 && ((uintptr_t)d->out_f32 & 15) == 0;
  const int resid_stages = (SMEM_BUDGET - NUM_EPI_WARPS * (STG_BYTES_PER_WARP_RESID - STG_BYTES_PER_WARP)) / p.stage_bytes;
  const int blocks_in_flight = resid_stages * (any_halo ? 3 : p.kb);      // 64-wide K blocks the ring holds with the larger staging
  const bool use_fast_resid = fast_resid && resid_stages >= 2 && blocks_in_flight >= 3;   // a starved operand ring costs more than the epilogue gains
  fast = fast && resid_stages >= 2 && blocks_in_flight >= 3;
  if (use_fast_resid || fast) {
    p.stg_warp_bytes = STG_BYTES_PER_WARP_RESID;
    if (resid_stages < p.stages) p.stages = resid_stages;
    // TMA-store mode of the bf16 row epilogue where the epilogue paces the tile: narrow N tile (the MMAs of a tile are short) with
    // the LayerNorm epilogue, or a short K loop.  Every warp's column share must be whole 128-byte panels.
    static const char* ts_env = getenv("PRV2_GEMM_TMA_STORE");      // diagnostics: 0 = never, 1 = wherever possible
    const long long mma_clk = (long long)(d->Ktot / BK) * (d->block_n > 128 ? 512 : 256);     // tensor-pipe clocks of one tile's K loop
    bool want = fast && d->block_n % 128 == 0 && mma_clk < (d->epi == PRV2_EPI_LN_GELU ? 12000 : 6000);
    if (ts_env && ts_env[0] == '0') want = false;
    if (ts_env && ts_env[0] == '1') want = fast && d->block_n % 128 == 0;
    if (want) {
      p.tma_store = 1;
      p.stg_warp_bytes = 8192;                                      // two 4 KB panels per warp
      const int box_w = d->tile_w >= 32 ? 32 : d->tile_w, box_h = 32 / box_w;
      cuuint64_t dims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
      cuuint64_t strides[3] = {(cuuint64_t)d->out_cs * 2, (cuuint64_t)d->W * d->out_cs * 2, (cuuint64_t)d->H * d->W * d->out_cs * 2};
      cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      CUresult r = enc(&p.tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)d->out_hi, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("prv2_umma_gemm: cuTensorMapEncodeTiled(out) failed (%d) Cout=%d cs=%d", (int)r, d->Cout, d->out_cs); return PRV2_ECUDA; }
    }
  }
  // always request the full budget: guarantees one CTA per SM, so a 512-column TMEM allocation can never deadlock
  int grid = p.total_tiles * cg < g_num_sms ? p.total_tiles * cg : g_num_sms;
  if (cg == 2) grid &= ~1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM_TOTAL; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  const int act = d->act;
  PRV2_CHECK_ARG(act >= PRV2_ACT_NONE && act <= PRV2_ACT_IDENTITY, "prv2_umma_gemm: unknown activation %d", act);
  PRV2_CHECK_ARG(act != PRV2_ACT_SIGMOID_GATE || (d->epi == PRV2_EPI_STORE && d->res_hi),
                 "prv2_umma_gemm: SIGMOID_GATE needs the STORE epilogue with `res` = the gated tensor (res2 is then added, relu copy allowed)");
  cudaError_t err = cudaSuccess;
#define PRV2_L2(E, A, F) (cg == 2 ? launch<E, A, 2, F>(cfg, p) : launch<E, A, 1, F>(cfg, p))
#define PRV2_L(E, A) (fast ? PRV2_L2(E, A, true) : PRV2_L2(E, A, false))
  switch (d->epi) {
    case PRV2_EPI_STORE:
      err = act == PRV2_ACT_RELU ? PRV2_L(PRV2_EPI_STORE, PRV2_ACT_RELU)
            : act == PRV2_ACT_GELU ? PRV2_L(PRV2_EPI_STORE, PRV2_ACT_GELU)
            : act == PRV2_ACT_GELU_TANH ? PRV2_L(PRV2_EPI_STORE, PRV2_ACT_GELU_TANH)
            : act == PRV2_ACT_SIGMOID_GATE ? PRV2_L2(PRV2_EPI_STORE, PRV2_ACT_SIGMOID_GATE, false)
                                        : PRV2_L(PRV2_EPI_STORE, PRV2_ACT_NONE);
      break;
    case PRV2_EPI_LN_GELU:
      err = act == PRV2_ACT_GELU_TANH ? PRV2_L(PRV2_EPI_LN_GELU, PRV2_ACT_GELU_TANH)
            : act == PRV2_ACT_RELU    ? PRV2_L(PRV2_EPI_LN_GELU, PRV2_ACT_RELU)
            : act == PRV2_ACT_IDENTITY ? PRV2_L(PRV2_EPI_LN_GELU, PRV2_ACT_NONE)         // LayerNorm only (SingleConvCNNLNHeavy)
                                      : PRV2_L(PRV2_EPI_LN_GELU, PRV2_ACT_GELU);
      break;
    case PRV2_EPI_RESID_F32: err = use_fast_resid ? PRV2_L2(PRV2_EPI_RESID_F32, PRV2_ACT_NONE, true) : PRV2_L2(PRV2_EPI_RESID_F32, PRV2_ACT_NONE, false); break;
    case PRV2_EPI_F32: err = PRV2_L2(PRV2_EPI_F32, PRV2_ACT_NONE, false); break;
    case PRV2_EPI_SHUFFLE: err = act == PRV2_ACT_RELU ? PRV2_L2(PRV2_EPI_SHUFFLE, PRV2_ACT_RELU, false) : PRV2_L2(PRV2_EPI_SHUFFLE, PRV2_ACT_NONE, false); break;
    default: err = PRV2_L2(PRV2_EPI_HEAD, PRV2_ACT_NONE, false); break;
  }
#undef PRV2_L
#undef PRV2_L2
  PRV2_CUDA(err);
  PRV2_LAUNCH_CHECK();
  return PRV2_OK;
}
