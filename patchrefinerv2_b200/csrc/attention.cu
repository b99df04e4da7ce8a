// Fused multi-head attention for the DINOv2 blocks (attention.py:49-62): softmax(q k^T / 8) v with
// head_dim 64, on tcgen05.  One CTA = 128 query rows of one (image, head).  Per 128-key chunk:
//   S = Q K^T      tcgen05.mma  (A = Q tile, B = K tile, both K-major SW128 straight from TMA)
//   online softmax 128 threads, thread t owns query row t = TMEM lane t (no shuffles)
//   O_c = P V      tcgen05.mma  (A = P written to smem as bf16 in the SW128 K-major layout,
//                                B = V tile as MN-major operand: rows are keys, as TMA delivers it)
//   acc = acc * alpha + O_c   in registers (fp32)
// The 1025-token sequence is 9 chunks; keys >= T are masked.  X3 = (hi,lo) operand splitting for
// the fp32-class precision mode (3 MMAs per product).
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <type_traits>

using namespace prv2;

namespace {

constexpr int TILE_BYTES = 128 * 64 * 2;   // 16 KB: 128 rows x 64 bf16
constexpr uint64_t SPIN_LIMIT_NS = 4000000000ull;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t globaltimer_ns() { uint64_t t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 0x3fff) == 0 && globaltimer_ns() - t0 > SPIN_LIMIT_NS) {
      printf("prv2_attention: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(adesc),
               "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
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
// both operand layouts use 128-byte rows, 8-row groups 1024 B apart, SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

struct alignas(64) AttnParams {
  CUtensorMap tm_hi, tm_lo;
  bf16* out_hi; bf16* out_lo;
  int B, T, heads, D;
};

template <bool X3>
__global__ void __launch_bounds__(128) attention_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  // tiles: Q, K, V, P0, P1 [, Ql, Kl, Vl, Pl0, Pl1]
  const uint32_t sQ = base, sK = base + TILE_BYTES, sV = base + 2 * TILE_BYTES, sP = base + 3 * TILE_BYTES;
  const uint32_t sQl = base + 5 * TILE_BYTES, sKl = base + 6 * TILE_BYTES, sVl = base + 7 * TILE_BYTES, sPl = base + 8 * TILE_BYTES;
  const uint32_t bars = base + (X3 ? 10 : 5) * TILE_BYTES;
  const uint32_t bar_q = bars, bar_kv = bars + 8, bar_s = bars + 16, bar_o = bars + 24, tmem_slot = bars + 32;
  uint8_t* pP = base_ptr + 3 * TILE_BYTES;
  uint8_t* pPl = base_ptr + 8 * TILE_BYTES;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128, head = blockIdx.y, b = blockIdx.z;
  const int D = p.D;
  const int n_chunks = (p.T + 127) / 128;

  if (tid == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_kv, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tS = tmem_base, tO = tmem_base + 128;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;

  const uint32_t kv_bytes = (X3 ? 4 : 2) * TILE_BYTES;
  if (tid == 0) {
    mbar_expect_tx(bar_q, (X3 ? 2 : 1) * TILE_BYTES);
    tma_load_3d(sQ, &p.tm_hi, bar_q, head * 64, q0, b);
    if (X3) tma_load_3d(sQl, &p.tm_lo, bar_q, head * 64, q0, b);
    mbar_expect_tx(bar_kv, kv_bytes);
    tma_load_3d(sK, &p.tm_hi, bar_kv, D + head * 64, 0, b);
    tma_load_3d(sV, &p.tm_hi, bar_kv, 2 * D + head * 64, 0, b);
    if (X3) {
      tma_load_3d(sKl, &p.tm_lo, bar_kv, D + head * 64, 0, b);
      tma_load_3d(sVl, &p.tm_lo, bar_kv, 2 * D + head * 64, 0, b);
    }
  }
  // instruction descriptors: c=F32, a=b=BF16, M=128; S: N=128 both K-major; PV: N=64, B MN-major (bit 16)
  const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
  const float c_log2 = 0.125f * 1.4426950408889634f;    // head_dim^-0.5 * log2(e)   (attention.py:41)

  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  mbar_wait(bar_q, 0);
  for (int j = 0; j < n_chunks; ++j) {
    const uint32_t ph = j & 1;
    mbar_wait(bar_kv, ph);
    if (tid == 0) {
      tc_fence_after();
      uint32_t accum = 0;
      for (int k = 0; k < 4; ++k) { tc_mma_bf16(tS, umma_desc_sw128(sQ) + 2 * k, umma_desc_sw128(sK) + 2 * k, idesc_s, accum); accum = 1; }
      if (X3) {
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tS, umma_desc_sw128(sQ) + 2 * k, umma_desc_sw128(sKl) + 2 * k, idesc_s, 1);
        for (int k = 0; k < 4; ++k) tc_mma_bf16(tS, umma_desc_sw128(sQl) + 2 * k, umma_desc_sw128(sK) + 2 * k, idesc_s, 1);
      }
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, ph);
    tc_fence_after();
    const int key0 = j * 128;
    const int n_valid = min(128, p.T - key0);
    float v[32];
    // pass 1: row max over this chunk
    float m_new = m_run;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      tc_ld32(tS + lane_off + c * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) if (c * 32 + i < n_valid) m_new = fmaxf(m_new, v[i]);
    }
    const float alpha = exp2f((m_run - m_new) * c_log2);
    // pass 2: p = exp2((s - m) * c), write P (bf16) to smem in the SW128 K-major layout
    float l_add = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      tc_ld32(tS + lane_off + c * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float pv = (c * 32 + i < n_valid) ? exp2f((v[i] - m_new) * c_log2) : 0.f;
        v[i] = pv;
        l_add += pv;
      }
      // keys c*32 .. c*32+31 of row tid: tile (c>>1), 16-byte chunks ((c&1)*4 + g), g = 0..3
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int chunk = (c & 1) * 4 + g;
        const uint32_t off = (uint32_t)(c >> 1) * TILE_BYTES + tid * 128 + ((chunk ^ (tid & 7)) << 4);
        bf16x8 hi8, lo8;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          hi8.v[e] = f2bf(v[g * 8 + e]);
          if (X3) lo8.v[e] = f2bf(v[g * 8 + e] - bf2f(hi8.v[e]));
        }
        *reinterpret_cast<bf16x8*>(pP + off) = hi8;
        if (X3) *reinterpret_cast<bf16x8*>(pPl + off) = lo8;
      }
    }
    l_run = l_run * alpha + l_add;
    m_run = m_new;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      uint32_t accum = 0;
      for (int k = 0; k < 8; ++k) {      // 8 slices of 16 keys: A advances 32 B inside tile (k>>2), B advances 16 rows (2048 B)
        const uint64_t ad = umma_desc_sw128(sP + (k >> 2) * TILE_BYTES) + 2 * (k & 3);
        tc_mma_bf16(tO, ad, umma_desc_sw128(sV + k * 2048), idesc_o, accum);
        accum = 1;
        if (X3) {
          tc_mma_bf16(tO, ad, umma_desc_sw128(sVl + k * 2048), idesc_o, 1);
          tc_mma_bf16(tO, umma_desc_sw128(sPl + (k >> 2) * TILE_BYTES) + 2 * (k & 3), umma_desc_sw128(sV + k * 2048), idesc_o, 1);
        }
      }
      tc_commit(bar_o);
    }
    mbar_wait(bar_o, ph);
    tc_fence_after();
    if (tid == 0 && j + 1 < n_chunks) {       // K/V/P buffers are free again: prefetch the next chunk under the accumulate
      mbar_expect_tx(bar_kv, kv_bytes);
      tma_load_3d(sK, &p.tm_hi, bar_kv, D + head * 64, key0 + 128, b);
      tma_load_3d(sV, &p.tm_hi, bar_kv, 2 * D + head * 64, key0 + 128, b);
      if (X3) {
        tma_load_3d(sKl, &p.tm_lo, bar_kv, D + head * 64, key0 + 128, b);
        tma_load_3d(sVl, &p.tm_lo, bar_kv, 2 * D + head * 64, key0 + 128, b);
      }
    }
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      tc_ld32(tO + lane_off + c * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[c * 32 + i] = acc[c * 32 + i] * alpha + v[i];
    }
  }
  const int q = q0 + tid;
  if (q < p.T) {
    const float inv = 1.0f / l_run;
    const size_t o = ((size_t)b * p.T + q) * D + head * 64;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float t[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] = acc[g * 8 + e] * inv;
      act_store8(p.out_hi, X3 ? p.out_lo : nullptr, o + g * 8, t);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// v3: warp-specialised, software-pipelined kernel for the one-pass bf16 mode.
//   warp 0        TMA producer: Q tiles once, then a 2-stage ring of K / V chunks
//   warp 1        MMA issuer (converged warp, one elected lane): S_w = Q_w K^T, O_w += P_w V
//   warps 2..5    softmax warpgroup 0 (query tile 0), warps 6..9 warpgroup 1 (query tile 1)
// One CTA owns TWO 128-query tiles of one (image, head): every K/V chunk is loaded once and used twice, and
// while one warpgroup runs its softmax the tensor core works for the other.
// * O_w accumulates in TMEM over all key chunks (tcgen05.mma accumulate); the row owner rescales it in
//   place (tcgen05.ld / st) only when its running maximum grows by more than 2^8 ("lazy rescale": the
//   exponentials are taken against a stale maximum otherwise, which is exact after the final 1/l).
// * TMEM reads of S are double-buffered in registers so the load of the next 32 columns overlaps the
//   exponentials of the current ones.
// * Key chunks are 128 wide except the LAST, which may be up to 144 wide (UMMA N = 16..144), so the 1025-token
//   DINOv2 sequence is 7 x 128 + 129 keys in 8 chunks instead of 9; the leftover query rows (T mod 128 <= 16)
//   go to a small SIMT kernel instead of a whole extra 128-row tile.
// * Output rows are staged in shared memory and leave with one bulk async copy per row.
// TMEM: S0 | S1 (160-col slots) | O0 | O1 (64 cols each).
// ---------------------------------------------------------------------------------------------
constexpr int KV_STAGES = 2;
constexpr int KV_ROWS = 144;
constexpr int KV_TILE_BYTES = KV_ROWS * 128;       // 18 KB
constexpr int S_COLS = 160;                         // TMEM columns reserved per S accumulator (32-aligned)
constexpr int V2_THREADS = 320;                    // SPLIT = 1: one softmax thread per query row
constexpr int V4_THREADS = 64 + 2 * 256;            // SPLIT = 2: two threads per query row (16 softmax warps, 4 per SMSP)
constexpr int XCH_BYTES = 2 * 2 * 2 * 128 * 4;      // row-statistics exchange between the two halves of a row: [tile][parity][half][row]
constexpr int TAIL_MAX = 16;                        // leftover rows / keys folded away from a full extra tile
constexpr int OUT_PITCH = 144;                      // staged output row: 64 bf16 + 16 B pad
constexpr float RESCALE_LOG2 = 8.0f;                // lazy rescale threshold (log2 units)

__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ float ex2_ftz(float x) {      // one MUFU.EX2; inputs are <= 8, flush-to-zero is what softmax wants
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
      "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
      "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
      "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
      "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
      "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

struct alignas(64) AttnParamsV2 {
  CUtensorMap tm_q, tm_kv;
  bf16* out_hi;
  int B, T, heads, D, n_chunks, last_width, tq_main;
};

// SPLIT = 2 ("v4"): the softmax of a 128-row query tile is shared by TWO warpgroups -- the warps (tile, half, quadrant) with the same
// quadrant may read the same TMEM lanes, so thread (half, row) owns columns [64*half, 64*half+64) of its row's scores, writes P
// tile `half` and rescales / emits O columns [32*half, 32*half+32).  The halves meet once per chunk (row maximum, through shared
// memory and a 256-thread named barrier) and once at the end (row sum).  v3 ran 2 softmax warps per SM sub-partition and was
// latency-bound there (issue slots 45 % busy, MUFU 42 %); four warps per sub-partition hide the TMEM-load / MUFU / barrier latencies.
// (18 warps put 5 on one sub-partition: 16384 / (5 * 32) caps the SPLIT = 2 kernel at 96 registers, so it walks its two
// 32-column pieces through ONE register buffer; the other warps of the sub-partition cover the TMEM-load latency instead.
// Keeping all 64 scores of a thread in registers between the two passes -- one TMEM read per chunk instead of two -- was
// measured SLOWER (342 vs 292 us per 27-patch call): at 96 registers it spills.)
template <int SPLIT>
__global__ void __launch_bounds__(64 + 256 * SPLIT, 1) attention_v3_kernel(const __grid_constant__ AttnParamsV2 p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  // tiles: Q0 Q1 | K[2] | V[2] (18 KB each) | P0(3) P1(3)
  const uint32_t sQ = base, sK = base + 2 * TILE_BYTES, sV = sK + KV_STAGES * KV_TILE_BYTES, sP = sV + KV_STAGES * KV_TILE_BYTES;
  const uint32_t bars = sP + 6 * TILE_BYTES;
  const uint32_t bar_q = bars;
  auto kv_full = [&](int s) { return bars + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bars + 8u * (1 + KV_STAGES + s); };
  auto s_full = [&](int w) { return bars + 8u * (1 + 2 * KV_STAGES + w); };
  auto p_full = [&](int w) { return bars + 8u * (3 + 2 * KV_STAGES + w); };
  auto o_done = [&](int w) { return bars + 8u * (5 + 2 * KV_STAGES + w); };
  const uint32_t tmem_slot = bars + 8u * (7 + 2 * KV_STAGES);
  float* const xch = reinterpret_cast<float*>(base_ptr + (bars - base) + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 256, head = blockIdx.y, b = blockIdx.z;
  const int D = p.D, T = p.T;
  const int n_chunks = p.n_chunks;
  const int n_wg = (q0 + 128 < p.tq_main) ? 2 : 1;

  if (tid == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int w = 0; w < 2; ++w) { mbar_init(s_full(w), 1); mbar_init(p_full(w), 128 * SPLIT); mbar_init(o_done(w), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  auto width_of = [&](int j) { return j == n_chunks - 1 ? p.last_width : 128; };

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (converged warp, elected lane issues)
    if (elect_one()) {
      mbar_expect_tx(bar_q, n_wg * TILE_BYTES);
      for (int w = 0; w < n_wg; ++w) tma_load_3d(sQ + w * TILE_BYTES, &p.tm_q, bar_q, head * 64, q0 + w * 128, b);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_chunks; ++j) {
      mbar_wait(kv_empty(stage), phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(kv_full(stage), 2 * KV_TILE_BYTES);
        tma_load_3d(sK + stage * KV_TILE_BYTES, &p.tm_kv, kv_full(stage), D + head * 64, j * 128, b);
        tma_load_3d(sV + stage * KV_TILE_BYTES, &p.tm_kv, kv_full(stage), 2 * D + head * 64, j * 128, b);
      }
      __syncwarp();
      if (++stage == KV_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer
    auto issue_s = [&](int w, int stage, int width) {          // S_w = Q_w K^T
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(width >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t qd = umma_desc_sw128(sQ + w * TILE_BYTES), kd = umma_desc_sw128(sK + stage * KV_TILE_BYTES);
      tc_mma_bf16(tmem_base + w * S_COLS, qd, kd, idesc, 0);
      tc_mma_bf16(tmem_base + w * S_COLS, qd + 2, kd + 2, idesc, 1);
      tc_mma_bf16(tmem_base + w * S_COLS, qd + 4, kd + 4, idesc, 1);
      tc_mma_bf16(tmem_base + w * S_COLS, qd + 6, kd + 6, idesc, 1);
      tc_commit(s_full(w));
    };
    const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    auto issue_o = [&](int w, int stage, int width, uint32_t accum) {     // O_w (+)= P_w V
      for (int k = 0; k < (width >> 4); ++k) {     // 16-key slices: A advances 32 B inside P tile (k>>2), B advances 16 key rows
        tc_mma_bf16(tmem_base + 2 * S_COLS + w * 64, umma_desc_sw128(sP + (3 * w + (k >> 2)) * TILE_BYTES) + 2 * (k & 3),
                    umma_desc_sw128(sV + stage * KV_TILE_BYTES + k * 2048), idesc_o, accum);
        accum = 1;
      }
      tc_commit(o_done(w));
    };
    mbar_wait(bar_q, 0);
    mbar_wait(kv_full(0), 0);
    tc_fence_after();
    if (elect_one()) for (int w = 0; w < n_wg; ++w) issue_s(w, 0, width_of(0));
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_chunks; ++j) {
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == KV_STAGES) { nstage = 0; nphase ^= 1; }
      const bool more = j + 1 < n_chunks;
      for (int w = 0; w < n_wg; ++w) {
        mbar_wait(p_full(w), j & 1);          // P_w(j) is in smem (and O_w rescaled); S_w(j) has been consumed
        if (more && w == 0) mbar_wait(kv_full(nstage), nphase);
        tc_fence_after();
        if (elect_one()) {
          issue_o(w, stage, width_of(j), j > 0 ? 1u : 0u);
          if (more) issue_s(w, nstage, width_of(j + 1));
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(kv_empty(stage));   // K_j / V_j are free once everything issued so far has retired
      __syncwarp();
      stage = nstage; phase = nphase;
    }
  } else {
    // ------------------------------------------------ softmax warpgroups
    const int w = (warp - 2) / (4 * SPLIT);
    const int half = SPLIT == 2 ? ((warp - 2) >> 2) & 1 : 0;
    constexpr int NP = 4 / SPLIT;                            // 32-column pieces of a full 128-key chunk per thread
    const int col_base = half * (128 / SPLIT);               // first score column this thread owns
    if (w < n_wg) {
      const int row = (warp & 3) * 32 + lane;               // TMEM lane == query row of this tile
      const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t tS = tmem_base + w * S_COLS + lane_off + col_base, tO = tmem_base + 2 * S_COLS + w * 64 + lane_off + half * 32;
      auto xslot = [&](int parity, int h) { return xch + (((w * 2 + parity) * 2 + h) * 128 + row); };
      auto pair_sync = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + w) : "memory"); };
      uint8_t* const pP = base_ptr + (sP - base) + 3 * w * TILE_BYTES;
      const float c_log2 = 0.125f * 1.4426950408889634f;    // head_dim^-0.5 * log2(e)   (attention.py:41)
      const float tau = RESCALE_LOG2 / c_log2;
      float m_run = -INFINITY, l_run = 0.f;
      uint32_t ra[32], rb[32];
      // (Forcing the two warpgroups to alternate in the MUFU-heavy section with named barriers was measured 16 % SLOWER:
      // one warpgroup alone is latency-bound, not MUFU-bound, so letting both run concurrently overlaps better.)
      // One key chunk of this warpgroup's softmax.  GENERAL = the last chunk (ragged: masked keys, 16..144 wide); every other
      // chunk runs the specialisation with compile-time width 128 and no per-element predicates (the generic code spent
      // ~320 of its ~1350 instructions per chunk on ISETP / FSEL masking).
      auto chunk_body = [&](auto general_tag, const int j) {
        constexpr bool GENERAL = decltype(general_tag)::value;
        const int n_valid = GENERAL ? T - j * 128 : 128;
        const int width = GENERAL ? p.last_width : 128;
        const int n_pieces = GENERAL ? max(1, min(NP, (min(width, 128) - col_base + 31) >> 5)) : NP;   // 32-column pieces of this thread inside the first 128 columns
        const bool wide = GENERAL && width > 128 && half == SPLIT - 1;     // keys 128..143 of the wide last chunk
        mbar_wait(s_full(w), j & 1);
        tc_fence_after();
        // ---- pass 1: row maximum (TMEM loads double-buffered in registers)
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};      // independent chains: 2 warps per SMSP hide little latency
        auto pmax = [&](const uint32_t (&r)[32], int col0) {
          if (!GENERAL || col0 + 32 <= n_valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (col0 + i < n_valid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(r[i]));
          }
        };
        tc_ld32_issue(tS, ra);
        tc_ld_wait();
        if (SPLIT == 2) {
          pmax(ra, col_base);
          if (n_pieces > 1) { tc_ld32_issue(tS + 32, ra); tc_ld_wait(); pmax(ra, col_base + 32); }
        } else {
          if (n_pieces > 1) tc_ld32_issue(tS + 32, rb);
          pmax(ra, 0);
          if (n_pieces > 1) {
            tc_ld_wait();
            if (n_pieces > 2) tc_ld32_issue(tS + 64, ra);
            pmax(rb, 32);
            if (n_pieces > 2) {
              tc_ld_wait();
              if (n_pieces > 3) tc_ld32_issue(tS + 96, rb);
              pmax(ra, 64);
              if (n_pieces > 3) { tc_ld_wait(); pmax(rb, 96); }
            }
          }
        }
        uint32_t rw[16];
        if (wide) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
              : "=r"(rw[0]), "=r"(rw[1]), "=r"(rw[2]), "=r"(rw[3]), "=r"(rw[4]), "=r"(rw[5]), "=r"(rw[6]), "=r"(rw[7]), "=r"(rw[8]), "=r"(rw[9]),
                "=r"(rw[10]), "=r"(rw[11]), "=r"(rw[12]), "=r"(rw[13]), "=r"(rw[14]), "=r"(rw[15])
              : "r"(tS - col_base + 128)
              : "memory");
          tc_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) if (128 + i < n_valid) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(rw[i]));
        }
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        if (SPLIT == 2) {                                     // the row's other half: slots alternate by chunk parity, so a fast
          *xslot(j & 1, half) = mx;                           // partner's write for chunk j+1 cannot overtake this chunk's read
          pair_sync();
          mx = fmaxf(mx, *xslot(j & 1, half ^ 1));
        }
        // ---- lazy rescale decision: keep the stale maximum unless the new one is > 2^8 larger
        const bool need = mx > m_run + tau;
        float alpha = 1.f;
        if (need) {
          alpha = ex2_ftz((m_run - mx) * c_log2);             // 0 on the first chunk (m_run = -inf)
          m_run = mx;
          l_run *= alpha;
        }
        const float mc = m_run * c_log2;
        // first piece of pass 2 can be fetched while we wait for the previous P V product (SPLIT == 1: it has the registers for it)
        if (SPLIT == 1) tc_ld32_issue(tS, ra);
        if (j > 0) {
          mbar_wait(o_done(w), (j - 1) & 1);                // P_w(j-1) V(j-1) has retired: P buffer and O_w are ours
          tc_fence_after();
          if (__any_sync(0xffffffffu, need)) {              // warp-uniform: tcgen05.ld / st are warp collectives
            if (SPLIT == 1) {
              tc_ld_wait();                                   // (drains the prefetched S piece too)
              uint32_t ro[32];
              tc_ld32_issue(tO, ro);
              tc_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
              tc_st32(tO, ro);
              tc_ld32_issue(tO + 32, ro);
              tc_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
              tc_st32(tO + 32, ro);
            } else {                                          // this half's 32 O columns, through the (still empty) S buffer
              tc_ld32_issue(tO, ra);
              tc_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) ra[i] = __float_as_uint(__uint_as_float(ra[i]) * alpha);
              tc_st32(tO, ra);
            }
            tc_st_wait();
          }
        }
        if (SPLIT == 2) tc_ld32_issue(tS, ra);
        // ---- pass 2: p = 2^(s*c - m*c) -> bf16 -> P tile (SW128 K-major), row sum
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        auto emit = [&](const uint32_t* r, int col0, int n) {
          const bool masked = GENERAL && col0 + n > n_valid;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g * 8 >= n) break;
            float e[8];
            if (!masked) {
#pragma unroll
              for (int t = 0; t < 8; ++t) e[t] = ex2_ftz(fmaf(__uint_as_float(r[g * 8 + t]), c_log2, -mc));
            } else {
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const float pv = ex2_ftz(fmaf(__uint_as_float(r[g * 8 + t]), c_log2, -mc));
                e[t] = (col0 + g * 8 + t < n_valid) ? pv : 0.f;
              }
            }
            l4[0] += e[0] + e[4]; l4[1] += e[1] + e[5]; l4[2] += e[2] + e[6]; l4[3] += e[3] + e[7];
            const int key = col0 + g * 8;                  // 8 consecutive keys = one 16-byte chunk of the P row
            const int tile = key >> 6, chunk = (key & 63) >> 3;
            __nv_bfloat162 h2[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) h2[t] = __floats2bfloat162_rn(e[2 * t], e[2 * t + 1]);
            *reinterpret_cast<uint4*>(pP + tile * TILE_BYTES + row * 128 + ((chunk ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(h2);
          }
        };
        if (wide) emit(rw, 128, 16);
        tc_ld_wait();
        if (SPLIT == 2) {
          emit(ra, col_base, 32);
          if (n_pieces > 1) { tc_ld32_issue(tS + 32, ra); tc_ld_wait(); emit(ra, col_base + 32, 32); }
        } else {
          if (n_pieces > 1) tc_ld32_issue(tS + 32, rb);
          emit(ra, 0, 32);
          if (n_pieces > 1) {
            tc_ld_wait();
            if (n_pieces > 2) tc_ld32_issue(tS + 64, ra);
            emit(rb, 32, 32);
            if (n_pieces > 2) {
              tc_ld_wait();
              if (n_pieces > 3) tc_ld32_issue(tS + 96, rb);
              emit(ra, 64, 32);
              if (n_pieces > 3) { tc_ld_wait(); emit(rb, 96, 32); }
            }
          }
        }
        l_run += (l4[0] + l4[1]) + (l4[2] + l4[3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // P rows -> visible to the tensor core
        tc_fence_before();
        mbar_arrive_local(p_full(w));
      };
      const bool last_general = p.last_width != 128 || T - (n_chunks - 1) * 128 < 128;
      for (int j = 0; j < n_chunks - 1; ++j) chunk_body(std::false_type{}, j);
      if (last_general) chunk_body(std::true_type{}, n_chunks - 1);
      else chunk_body(std::false_type{}, n_chunks - 1);
      // ---- epilogue: O / l -> bf16 -> staged row -> one bulk copy
      mbar_wait(o_done(w), (n_chunks - 1) & 1);
      tc_fence_after();
      if (SPLIT == 2) {                                       // row sum = the two halves' partial sums (same rescale history)
        *xslot(n_chunks & 1, half) = l_run;
        pair_sync();
        l_run += *xslot(n_chunks & 1, half ^ 1);
      }
      const float inv = 1.0f / l_run;
      uint8_t* const srow = pP + row * OUT_PITCH;             // the P tiles are free now
      tc_ld32_issue(tO, ra);
      if (SPLIT == 1) tc_ld32_issue(tO + 32, rb);
      tc_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        __nv_bfloat162 h2[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) h2[t] = __floats2bfloat162_rn(__uint_as_float(ra[g * 8 + 2 * t]) * inv, __uint_as_float(ra[g * 8 + 2 * t + 1]) * inv);
        *reinterpret_cast<uint4*>(srow + half * 64 + g * 16) = *reinterpret_cast<const uint4*>(h2);
        if (SPLIT == 1) {
#pragma unroll
          for (int t = 0; t < 4; ++t) h2[t] = __floats2bfloat162_rn(__uint_as_float(rb[g * 8 + 2 * t]) * inv, __uint_as_float(rb[g * 8 + 2 * t + 1]) * inv);
          *reinterpret_cast<uint4*>(srow + 64 + g * 16) = *reinterpret_cast<const uint4*>(h2);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (SPLIT == 2) pair_sync();                            // both halves of the staged row are written and fenced
      const int q = q0 + w * 128 + row;
      if (q < p.tq_main && half == 0) {
        bf16* const gdst = p.out_hi + ((size_t)b * T + q) * D + head * 64;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(gdst), "r"(smem_u32(srow)) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Leftover query rows [tq_main, T) (1 row for the 1025-token DINOv2 sequence): one 256-thread CTA per
// (image, head, row).  Phase 1: thread-per-key scores into shared memory + block softmax statistics;
// phase 2: thread (kgroup, 4 dims) accumulates p*V over its keys, 16 key groups reduced through smem.
constexpr int TAIL_THREADS = 256;
constexpr int TAIL_MAX_T = 2048;
__global__ void __launch_bounds__(TAIL_THREADS) attention_tail_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int B, int T, int heads,
                                                                      int tq_main) {
  __shared__ float s_p[TAIL_MAX_T];
  __shared__ float s_q[64];
  __shared__ float s_red[TAIL_THREADS / 32];
  __shared__ float s_acc[16][64];
  const int n_tail = T - tq_main;
  const int r = blockIdx.x % n_tail, head = (blockIdx.x / n_tail) % heads, b = blockIdx.x / (n_tail * heads);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = heads * 64;
  const size_t row_stride = (size_t)3 * D;
  const bf16* base = qkv + (size_t)b * T * row_stride;
  if (tid < 64) s_q[tid] = bf2f(base[(size_t)(tq_main + r) * row_stride + head * 64 + tid]);
  __syncthreads();
  const float c_log2 = 0.125f * 1.4426950408889634f;
  // phase 1: scores, thread per key (an 8-lanes-per-key "coalesced" variant with shuffle reduction measured 2x slower:
  // the K rows are L2 hits and the extra passes cost more than the uncoalesced lines)
  float m = -INFINITY;
  for (int k = tid; k < T; k += TAIL_THREADS) {
    const bf16* kp = base + (size_t)k * row_stride + D + head * 64;
    float sc = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float t[8];
      act_load8(kp, nullptr, g * 8, t);
#pragma unroll
      for (int e = 0; e < 8; ++e) sc = fmaf(s_q[g * 8 + e], t[e], sc);
    }
    s_p[k] = sc;
    m = fmaxf(m, sc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  m = s_red[0];
#pragma unroll
  for (int w = 1; w < TAIL_THREADS / 32; ++w) m = fmaxf(m, s_red[w]);
  __syncthreads();
  float l = 0.f;
  for (int k = tid; k < T; k += TAIL_THREADS) {
    const float pv = exp2f((s_p[k] - m) * c_log2);
    s_p[k] = pv;
    l += pv;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if (lane == 0) s_red[warp] = l;
  __syncthreads();
  l = 0.f;
#pragma unroll
  for (int w = 0; w < TAIL_THREADS / 32; ++w) l += s_red[w];
  // phase 2: out[d] = sum_k p_k V[k, d]
  const int kg = tid >> 4, d4 = (tid & 15) * 4;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const bf16* vbase = base + 2 * D + head * 64 + d4;
  int k = kg;
  for (; k + 7 * 16 < T; k += 8 * 16) {                      // eight independent 8-byte loads in flight per thread
    uint2 raw[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) raw[u] = *reinterpret_cast<const uint2*>(vbase + (size_t)(k + u * 16) * row_stride);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
      const float2 f0 = __bfloat1622float2(v2[0]), f1 = __bfloat1622float2(v2[1]);
      const float pv = s_p[k + u * 16];
      a0 = fmaf(pv, f0.x, a0); a1 = fmaf(pv, f0.y, a1); a2 = fmaf(pv, f1.x, a2); a3 = fmaf(pv, f1.y, a3);
    }
  }
  for (; k < T; k += 16) {
    const uint2 raw = *reinterpret_cast<const uint2*>(vbase + (size_t)k * row_stride);
    const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
    const float2 f0 = __bfloat1622float2(v2[0]), f1 = __bfloat1622float2(v2[1]);
    const float pv = s_p[k];
    a0 = fmaf(pv, f0.x, a0); a1 = fmaf(pv, f0.y, a1); a2 = fmaf(pv, f1.x, a2); a3 = fmaf(pv, f1.y, a3);
  }
  s_acc[kg][d4] = a0; s_acc[kg][d4 + 1] = a1; s_acc[kg][d4 + 2] = a2; s_acc[kg][d4 + 3] = a3;
  __syncthreads();
  if (tid < 64) {
    float a = 0.f;
#pragma unroll
    for (int g = 0; g < 16; ++g) a += s_acc[g][tid];
    out[((size_t)b * T + tq_main + r) * D + head * 64 + tid] = f2bf(a / l);
  }
}

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

bool g_attr_set = false;

}  // namespace

extern "C" int prv2_attention(const prv2_bf16* qkv_hi, const prv2_bf16* qkv_lo, int B, int T, int heads, prv2_bf16* out_hi, prv2_bf16* out_lo,
                              prv2_stream_t stream) {
  PRV2_CHECK_ARG(qkv_hi && out_hi, "prv2_attention: null pointer");
  PRV2_CHECK_ARG((qkv_lo == nullptr) == (out_lo == nullptr), "prv2_attention: lo planes must both be present or absent");
  PRV2_CHECK_ARG(B > 0 && T > 0 && heads > 0 && heads <= 65535 && B <= 65535, "prv2_attention: bad shape");
  auto enc = get_encode();
  if (!enc) { set_error("prv2_attention: cuTensorMapEncodeTiled unavailable"); return PRV2_ECUDA; }
  const int D = heads * 64;
  AttnParams p;
  memset(&p, 0, sizeof(p));
  cuuint64_t dims[3] = {(cuuint64_t)3 * D, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)3 * D * 2, (cuuint64_t)T * 3 * D * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&p.tm_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)qkv_hi, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("prv2_attention: cuTensorMapEncodeTiled failed (%d)", (int)r); return PRV2_ECUDA; }
  p.tm_lo = p.tm_hi;
  if (qkv_lo) {
    r = enc(&p.tm_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)qkv_lo, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("prv2_attention: cuTensorMapEncodeTiled(lo) failed (%d)", (int)r); return PRV2_ECUDA; }
  }
  p.out_hi = (bf16*)out_hi; p.out_lo = (bf16*)out_lo;
  p.B = B; p.T = T; p.heads = heads; p.D = D;
  const int smem1 = 5 * TILE_BYTES + 1024 + 64, smem3 = 10 * TILE_BYTES + 1024 + 64;
  const int smem_v2 = 2 * TILE_BYTES + 2 * KV_STAGES * KV_TILE_BYTES + 6 * TILE_BYTES + 1024 + 256 + XCH_BYTES;
  static const char* split_env = getenv("PRV2_ATTN_SPLIT");            // diagnostics: 1 = one softmax thread per row (v3)
  const bool split2 = !(split_env && split_env[0] == '1');
  static const bool force_v1 = getenv("PRV2_ATTN_V1") != nullptr;
  if (!g_attr_set) {
    PRV2_CUDA(cudaFuncSetAttribute(attention_v3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_v2));
    PRV2_CUDA(cudaFuncSetAttribute(attention_v3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_v2));
    PRV2_CUDA(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    PRV2_CUDA(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
    g_attr_set = true;
  }
  dim3 grid(cdiv(T, 128), heads, B);
  if (qkv_lo) { attention_kernel<true><<<grid, 128, smem3, (cudaStream_t)stream>>>(p); PRV2_LAUNCH_CHECK(); return PRV2_OK; }
  if (force_v1) { attention_kernel<false><<<grid, 128, smem1, (cudaStream_t)stream>>>(p); PRV2_LAUNCH_CHECK(); return PRV2_OK; }

  AttnParamsV2 p2;
  memset(&p2, 0, sizeof(p2));
  p2.tm_q = p.tm_hi;
  cuuint32_t box_kv[3] = {64, KV_ROWS, 1};
  r = enc(&p2.tm_kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)qkv_hi, dims, strides, box_kv, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("prv2_attention: cuTensorMapEncodeTiled(kv) failed (%d)", (int)r); return PRV2_ECUDA; }
  p2.out_hi = (bf16*)out_hi;
  p2.B = B; p2.T = T; p2.heads = heads; p2.D = D;
  // key chunks: 128 wide, the last one 16..144 (multiple of 16) so that T mod 128 <= 16 does not cost a whole chunk
  p2.n_chunks = T <= KV_ROWS ? 1 : cdiv(T - TAIL_MAX, 128);
  p2.last_width = ((T - 128 * (p2.n_chunks - 1)) + 15) / 16 * 16;
  // query rows: leftover rows (<= 16) go to the SIMT tail kernel instead of a mostly empty 128-row tile
  const int rem = T % 128;
  p2.tq_main = (rem != 0 && rem <= TAIL_MAX && T > 128 && T <= TAIL_MAX_T) ? T - rem : T;
  if (split2) attention_v3_kernel<2><<<dim3(cdiv(p2.tq_main, 256), heads, B), V4_THREADS, smem_v2, (cudaStream_t)stream>>>(p2);
  else attention_v3_kernel<1><<<dim3(cdiv(p2.tq_main, 256), heads, B), V2_THREADS, smem_v2, (cudaStream_t)stream>>>(p2);
  PRV2_LAUNCH_CHECK();
  if (p2.tq_main < T) {
    attention_tail_kernel<<<B * heads * (T - p2.tq_main), TAIL_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)qkv_hi, (bf16*)out_hi, B, T, heads, p2.tq_main);
    PRV2_LAUNCH_CHECK();
  }
  return PRV2_OK;
}
