// Fused multi-head attention for the DINOv2 blocks (attention.py:49-62): softmax(q k^T / 8) v with head_dim 64, on tcgen05:
// S = Q K^T and O = P V on the tensor cores (S, P and O live in tensor memory), online softmax with one thread per query row.
// The design notes are at the kernel (attention_v6_kernel).  X3 = (hi, lo) operand splitting for the fp32-class precision mode
// (3 MMAs per product).
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <type_traits>
#include <mutex>

using namespace prv2;

namespace {

constexpr int TILE_BYTES = 128 * 64 * 2;   // 16 KB: 128 rows x 64 bf16
constexpr long long SPIN_LIMIT_CLK = 8000000000ll;     // ~4 s of SM clocks: a deadlock traps instead of hanging the GPU

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
          printf("prv2_attention: mbarrier wait timed out (block %d,%d,%d warp %d, barrier @%u parity %u)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x >> 5,
                 bar & 0x3ffu, parity);
        __trap();
      }
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
// both operand layouts use 128-byte rows, 8-row groups 1024 B apart, SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

constexpr int TAIL_MAX = 16;                        // leftover rows / keys folded away from a full extra tile
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

// ---------------------------------------------------------------------------------------------
// v6: one softmax THREAD per query row, ONE score slot per tile with P written over the scores, several independent CTAs per SM.
//   warps 0..3    softmax warpgroup of query tile 0 (thread = TMEM lane = query row: no shuffles, no shared-memory exchange)
//   warps 4..7    softmax warpgroup of query tile 1 ((hi, lo) mode only)
//   next warp     MMA issuer (one elected lane): S_w = Q_w K^T (operands in shared memory), O_w += P_w V (A = P_w from TMEM)
//   next warp     TMA producer: Q tiles once, then a ring of K / V chunks      ((hi, lo) mode: two more idle warps complete the warpgroup)
// What round 2's clock64 traces showed (gpurun_out/att_trace_*.log):
//   * a softmax warp cannot overlap its own MUFU, FMA and ALU work: the 64 exponentials of a 64-key chunk take ~960 clocks when the
//     warp has its scheduler to itself (512 clocks of MUFU), and ~600 more clocks per chunk go to synchronisation (barrier wait, TMEM
//     load / store round trips, arrive).  Pipes only overlap ACROSS warps;
//   * the chain softmax(j) -> P V(j), S(j+1) -> softmax(j+1) of one query tile is serial whatever the buffering (the MMA warp needs
//     ~700 clocks to wake up, issue eight MMAs and see them commit), so with two tiles per SM every pipe idled half the time; wider
//     chunks (128 keys), two score slots, or two threads per row all landed on the same ~290 us per 27-patch call.
// Hence the one-pass configuration: 64-key chunks, S 64 + O 64 = 128 TMEM columns and 50 KB of shared memory per CTA, FOUR CTAs per
// SM: four independent chains keep four softmax warps on every scheduler (222 us, 1.31x; MUFU pipe ~70 % busy).  The scores are read
// from tensor memory twice in 32-column pieces (pass 1: row maximum; pass 2: exponentials), so a thread holds 32 score registers (80
// registers per thread); P (16-bit pairs) overwrites the score columns pass 2 has already consumed, and the tensor pipe executes in
// issue order, so S_w(j+1) -- issued right after P_w(j) V(j) -- may overwrite P_w(j) safely.  POLY of every 8 exponentials are
// evaluated on the FMA pipe (Cody-Waite reduction x = n + f, |f| <= 1/2, cubic minimax polynomial for 2^f with relative error
// 7.5e-5 -- far below the 2^-9 of the bf16 P -- and the exponent inserted with an integer add: the FlashAttention-4 trick; a sweep
// of 0 / 2 / 4 / 6 of 8 gave 226 / 222 / 229 / 244 us: the kernel is no longer MUFU-bound alone, issue slots are as scarce).
// The row owner rescales O_w in place, lazily (only when the running maximum grew by more than 2^8), after P_w(j-1) V(j-1) retired.
// X3 = fp32-class mode: Q, K, V arrive as (hi, lo) FP16 planes, S = Qh Kl + Ql Kh + Qh Kh, P is split into (hi, lo) planes (P_lo in
// its own 64 columns), O = Ph Vl + Pl Vh + Ph Vh; 128-key chunks, two tiles per CTA, one CTA per SM (tensor-bound: 3 MMAs per
// product), piece loads software-pipelined, all exponentials on the MUFU pipe; the output is written as (hi, lo) planes.
// The last key chunk holds the remaining 1..CH keys; its MMAs run 16-key slices up to the next multiple of 16 (TMA zero-fills rows
// past T), so the 1025th DINOv2 token costs one 16-wide slice.  Leftover query rows (T mod 128 <= 16) go to the SIMT tail kernel.
// TMEM per tile: S [0, CH) (P_hi over its first half) | X3: P_lo [128, 192) | O (64 columns).
// ---------------------------------------------------------------------------------------------
constexpr float TRUNC_BIAS_LOG2 = 0.0028150156f;    // log2(1 + 2^-9)
#ifndef PRV2_ATTN_POLY
#define PRV2_ATTN_POLY 2                            // exponentials per 8 evaluated on the FMA pipe (one-pass mode), 0..8, even
#endif

__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem_d), "r"(tmem_a),
               "l"(bdesc), "r"(idesc), "r"(accumulate)
               : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
      "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// two independent fp32 operations per instruction (FFMA2 / FADD2 on sm_100)
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b));
  return v;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2v(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2v(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

struct alignas(64) AttnParamsV6 {
  CUtensorMap tm_hi, tm_lo;       // box {64 ch, 128 tokens, 1 image} over the qkv planes: Q tiles
  CUtensorMap tmkv_hi, tmkv_lo;   // box {64 ch, CH tokens, 1 image}: K / V chunks
  bf16* out_hi; bf16* out_lo;
  int B, T, heads, D, n_chunks, last_valid, tq_main;
  unsigned long long* trace;      // diagnostics (PRV2_ATTN_TRACE=1): clock64 stamps of CTA (0,0,0), [role 0..2][chunk][5]
};
// (compiled in only with -DPRV2_ATTN_TRACE_BUILD: the stamps cost ~10 instructions each in the MMA and softmax loops)
#ifdef PRV2_ATTN_TRACE_BUILD
#define TRACE(role, j, k)                                                                                             \
  do {                                                                                                                \
    if (p.trace && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (j) < 32) p.trace[((role) * 32 + (j)) * 5 + (k)] = clock64(); \
  } while (0)
#else
#define TRACE(role, j, k) do { } while (0)
#endif

// Leftover query rows [tq_main, T) (1 row for the 1025-token DINOv2 sequence) are computed in plain SIMT code by a small second
// kernel, one 256-thread CTA per (row, head, image): a 128-row tensor-core tile for one row would cost an eighth of the whole
// attention.  (Folding these rows into the main launch as extra CTAs was measured slower: a latency-bound SIMT CTA holds one of
// the two 97 KB CTA slots of an SM for longer than a tile CTA.)  Phase 1: thread-per-key scores into shared memory + block softmax
// statistics; phase 2: thread (key group, 4 dims) accumulates p*V over its keys, the key groups are reduced through smem.
// lo == nullptr: plain bf16 planes; otherwise value = hi + lo (fp32-class mode) and the output is split the same way.
constexpr int TAIL_MAX_T = 2048;
template <int NTHR>
__device__ __forceinline__ void attention_tail_row(float* smem, const bf16* __restrict__ qkv, const bf16* __restrict__ qkv_lo, bf16* __restrict__ out,
                                                   bf16* __restrict__ out_lo, int T, int heads, int b, int head, int qrow) {
  constexpr int KG = NTHR / 16;
  float* s_p = smem;                         // [TAIL_MAX_T]
  float* s_q = s_p + TAIL_MAX_T;             // [64]
  float* s_red = s_q + 64;                   // [NTHR / 32]
  float* s_acc = s_red + 32;                 // [KG][64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = heads * 64;
  const size_t row_stride = (size_t)3 * D;
  const size_t img = (size_t)b * T * row_stride;
  const bf16* base = qkv + img;
  const bf16* base_lo = qkv_lo ? qkv_lo + img : nullptr;
  __syncthreads();                            // (the previous row's reads of the shared arrays are done)
  if (tid < 64) s_q[tid] = act_load(base, base_lo, (size_t)qrow * row_stride + head * 64 + tid);
  __syncthreads();
  const float c_log2 = 0.125f * 1.4426950408889634f;
  // phase 1: scores, thread per key (an 8-lanes-per-key "coalesced" variant with shuffle reduction measured 2x slower:
  // the K rows are L2 hits and the extra passes cost more than the uncoalesced lines)
  float m = -INFINITY;
  for (int k = tid; k < T; k += NTHR) {
    const size_t ko = (size_t)k * row_stride + D + head * 64;
    float sc = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float t[8];
      act_load8(base + ko, base_lo ? base_lo + ko : nullptr, g * 8, t);
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
  for (int w = 1; w < NTHR / 32; ++w) m = fmaxf(m, s_red[w]);
  __syncthreads();
  float l = 0.f;
  for (int k = tid; k < T; k += NTHR) {
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
  for (int w = 0; w < NTHR / 32; ++w) l += s_red[w];
  // phase 2: out[d] = sum_k p_k V[k, d]
  const int kg = tid >> 4, d4 = (tid & 15) * 4;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const size_t voff = 2 * D + head * 64 + d4;
  auto load4 = [&](const bf16* pl, int k, float (&f)[4], bool add) {
    const uint2 raw = *reinterpret_cast<const uint2*>(pl + voff + (size_t)k * row_stride);
    float2 f0, f1;
    if (base_lo) {                            // (hi, lo) planes are FP16
      const __half2* v2 = reinterpret_cast<const __half2*>(&raw);
      f0 = __half22float2(v2[0]); f1 = __half22float2(v2[1]);
    } else {
      const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
      f0 = __bfloat1622float2(v2[0]); f1 = __bfloat1622float2(v2[1]);
    }
    if (add) { f[0] += f0.x; f[1] += f0.y; f[2] += f1.x; f[3] += f1.y; }
    else { f[0] = f0.x; f[1] = f0.y; f[2] = f1.x; f[3] = f1.y; }
  };
  int k = kg;
  for (; k + 7 * KG < T; k += 8 * KG) {                      // eight independent 8-byte loads in flight per thread (and plane)
    float f[8][4];
#pragma unroll
    for (int u = 0; u < 8; ++u) load4(base, k + u * KG, f[u], false);
    if (base_lo) {
#pragma unroll
      for (int u = 0; u < 8; ++u) load4(base_lo, k + u * KG, f[u], true);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float pv = s_p[k + u * KG];
      a0 = fmaf(pv, f[u][0], a0); a1 = fmaf(pv, f[u][1], a1); a2 = fmaf(pv, f[u][2], a2); a3 = fmaf(pv, f[u][3], a3);
    }
  }
  for (; k < T; k += KG) {
    float f[4];
    load4(base, k, f, false);
    if (base_lo) load4(base_lo, k, f, true);
    const float pv = s_p[k];
    a0 = fmaf(pv, f[0], a0); a1 = fmaf(pv, f[1], a1); a2 = fmaf(pv, f[2], a2); a3 = fmaf(pv, f[3], a3);
  }
  s_acc[kg * 64 + d4] = a0; s_acc[kg * 64 + d4 + 1] = a1; s_acc[kg * 64 + d4 + 2] = a2; s_acc[kg * 64 + d4 + 3] = a3;
  __syncthreads();
  if (tid < 64) {
    float a = 0.f;
#pragma unroll
    for (int g = 0; g < KG; ++g) a += s_acc[g * 64 + tid];
    act_store(out, out_lo, ((size_t)b * T + qrow) * D + head * 64 + tid, a / l);
  }
}

__global__ void __launch_bounds__(256) attention_tail_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ qkv_lo, bf16* __restrict__ out,
                                                             bf16* __restrict__ out_lo, int T, int heads, int tq_main) {
  __shared__ float smem[TAIL_MAX_T + 64 + 32 + 16 * 64];
  attention_tail_row<256>(smem, qkv, qkv_lo, out, out_lo, T, heads, blockIdx.z, blockIdx.y, tq_main + blockIdx.x);
}


// NT = query tiles per CTA, CTAS = CTAs per SM.  One-pass mode: NT = 1, four CTAs per SM (128 TMEM columns, 50 KB of shared memory,
// 192 threads x 80 registers each).  (hi, lo) mode: NT = 2, one CTA per SM (512 TMEM columns, 193 KB, setmaxnreg 192 / 112).
template <bool X3> struct V6Cfg {
  static constexpr int NT = X3 ? 2 : 1;
  static constexpr int CTAS = X3 ? 1 : 4;                          // CTAs per SM
  static constexpr int CH = X3 ? 128 : 64;                         // keys per chunk
  static constexpr int KVB = CH * 128;                             // bytes of one K or V chunk buffer
  static constexpr int PLANES = X3 ? 2 : 1;
  static constexpr int STAGES = 2;
  static constexpr int STAGE_BYTES = 2 * PLANES * KVB;             // K_hi | V_hi [| K_lo | V_lo]
  static constexpr int SMEM = NT * PLANES * TILE_BYTES + STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int THREADS = X3 ? 128 * NT + 128 : 192;        // softmax warpgroup(s) + the MMA and TMA warps (X3: a whole warpgroup, for setmaxnreg)
  static constexpr int TILE_COLS = X3 ? 256 : 128;                 // S CH (P_hi over its first half) | X3: P_lo 64 | O 64
  static constexpr int P_LO_COL = 128;
  static constexpr int O_COL = X3 ? 192 : 64;
  static constexpr int TMEM_COLS = X3 ? 512 : 128;
};

template <bool X3, int POLY_T>
__global__ void __launch_bounds__(V6Cfg<X3>::THREADS, V6Cfg<X3>::CTAS) attention_v6_kernel(const __grid_constant__ AttnParamsV6 p) {
  constexpr int PLANES = V6Cfg<X3>::PLANES, STAGES = V6Cfg<X3>::STAGES, STAGE_BYTES = V6Cfg<X3>::STAGE_BYTES, NT = V6Cfg<X3>::NT;
  constexpr int CH = V6Cfg<X3>::CH, KVB = V6Cfg<X3>::KVB, P_LO_COL = V6Cfg<X3>::P_LO_COL;
  constexpr int O_COL = V6Cfg<X3>::O_COL, TILE_COLS = V6Cfg<X3>::TILE_COLS, MMA_WARP = 4 * NT, TMA_WARP = MMA_WARP + 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  // Q tiles: [plane][tile] 16 KB each | K/V ring
  const uint32_t sQ = base, sRing = base + NT * PLANES * TILE_BYTES;
  const uint32_t bars = sRing + STAGES * STAGE_BYTES;
  const uint32_t bar_q = bars;
  auto kv_full = [&](int s) { return bars + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bars + 8u * (1 + STAGES + s); };
  auto s_full = [&](int w) { return bars + 8u * (1 + 2 * STAGES + w); };
  auto p_full = [&](int w) { return bars + 8u * (3 + 2 * STAGES + w); };
  auto o_done = [&](int w) { return bars + 8u * (5 + 2 * STAGES + w); };
  auto o_final = [&](int w) { return bars + 8u * (7 + 2 * STAGES + w); };
  const uint32_t tmem_slot = bars + 8u * (9 + 2 * STAGES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * (128 * NT), head = blockIdx.y, b = blockIdx.z;
  const int D = p.D, T = p.T;
  const int n_chunks = p.n_chunks;
  const int n_wg = (NT == 2 && q0 + 128 < p.tq_main) ? 2 : 1;
  const int last_width = (p.last_valid + 15) & ~15;                 // columns the MMAs of the last chunk write / read

  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
  const long long cta_clk0 = clock64();
  if (p.trace && tid == 0 && cta_lin < 8192) p.trace[480 + 3 * cta_lin] = globaltimer_ns();
  if (tid == 0) {
    mbar_init(bar_q, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int w = 0; w < NT; ++w) {
      mbar_init(s_full(w), 1);
      mbar_init(p_full(w), 4);
      mbar_init(o_done(w), 1);
      mbar_init(o_final(w), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)V6Cfg<X3>::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp >= MMA_WARP) {
  // (the register split must dominate each role's code, or ptxas applies the minimum to all of it)
  if (X3) asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
  if (warp == TMA_WARP) {
    // ------------------------------------------------ TMA producer (converged warp, elected lane issues)
    if (elect_one()) {
      mbar_expect_tx(bar_q, n_wg * PLANES * TILE_BYTES);
      for (int w = 0; w < n_wg; ++w) {
        tma_load_3d(sQ + w * TILE_BYTES, &p.tm_hi, bar_q, head * 64, q0 + w * 128, b);
        if (X3) tma_load_3d(sQ + (NT + w) * TILE_BYTES, &p.tm_lo, bar_q, head * 64, q0 + w * 128, b);
      }
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_chunks; ++j) {
      mbar_wait(kv_empty(stage), phase ^ 1);
      if (elect_one()) {
        const uint32_t dst = sRing + stage * STAGE_BYTES;
        mbar_expect_tx(kv_full(stage), 2 * PLANES * KVB);          // rows past T are zero-filled by TMA and count like the others
        tma_load_3d(dst, &p.tmkv_hi, kv_full(stage), D + head * 64, j * CH, b);
        tma_load_3d(dst + KVB, &p.tmkv_hi, kv_full(stage), 2 * D + head * 64, j * CH, b);
        if (X3) {
          tma_load_3d(dst + 2 * KVB, &p.tmkv_lo, kv_full(stage), D + head * 64, j * CH, b);
          tma_load_3d(dst + 3 * KVB, &p.tmkv_lo, kv_full(stage), 2 * D + head * 64, j * CH, b);
        }
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------ MMA issuer
    // Everything the elected lane needs is warp-uniform and computed by the converged warp; the issue blocks are straight-line
    // (descriptor + constant) so the operands stay in uniform registers.
    constexpr uint32_t FMT = X3 ? 0u : 1u;              // operand format: F16 for the (hi, lo) pair planes, BF16 in one-pass mode
    const uint32_t idesc_s = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(CH >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_s_last = (idesc_s & ~(0x3fu << 17)) | ((uint32_t)(last_width >> 3) << 17);
    const uint32_t idesc_o = (1u << 4) | (FMT << 7) | (FMT << 10) | (1u << 16) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    // The tensor core adds every MMA into the fp32 accumulator with round-toward-zero (a bias of ~2^-24 of the running sum per
    // MMA, scripts/diag_accum.py), so in (hi, lo) mode the two small cross terms go FIRST and the main term last.
    auto issue_s = [&](uint32_t tS, uint64_t qh, uint64_t ql, uint64_t kh, uint64_t kl, uint32_t idesc, uint32_t bar) {   // S = Q K^T (X3: Qh Kl + Ql Kh + Qh Kh)
      if (X3) {
        tc_mma_bf16(tS, qh, kl, idesc, 0u);
        tc_mma_bf16(tS, qh + 2, kl + 2, idesc, 1u);
        tc_mma_bf16(tS, qh + 4, kl + 4, idesc, 1u);
        tc_mma_bf16(tS, qh + 6, kl + 6, idesc, 1u);
        tc_mma_bf16(tS, ql, kh, idesc, 1u);
        tc_mma_bf16(tS, ql + 2, kh + 2, idesc, 1u);
        tc_mma_bf16(tS, ql + 4, kh + 4, idesc, 1u);
        tc_mma_bf16(tS, ql + 6, kh + 6, idesc, 1u);
      }
      tc_mma_bf16(tS, qh, kh, idesc, X3 ? 1u : 0u);
      tc_mma_bf16(tS, qh + 2, kh + 2, idesc, 1u);
      tc_mma_bf16(tS, qh + 4, kh + 4, idesc, 1u);
      tc_mma_bf16(tS, qh + 6, kh + 6, idesc, 1u);
      tc_commit(bar);
    };
    // O (+)= P V, A = P from TMEM (X3: Ph Vl + Pl Vh + Ph Vh).  16-key slices: A advances 8 TMEM columns, B 16 key rows (2048 B = +128 in the descriptor)
    auto issue_o_slice = [&](uint32_t tO, uint32_t tP, uint64_t vh, uint64_t vl, int k, uint32_t accum) {
      if (X3) {
        tc_mma_bf16_ts(tO, tP + 8 * k, vl + 128 * k, idesc_o, accum);
        tc_mma_bf16_ts(tO, tP + P_LO_COL + 8 * k, vh + 128 * k, idesc_o, 1u);
      }
      tc_mma_bf16_ts(tO, tP + 8 * k, vh + 128 * k, idesc_o, X3 ? 1u : accum);
    };
    const uint64_t qd_h[2] = {umma_desc_sw128(sQ), umma_desc_sw128(sQ + TILE_BYTES)};
    const uint64_t qd_l[2] = {umma_desc_sw128(sQ + NT * TILE_BYTES), umma_desc_sw128(sQ + (NT + 1) * TILE_BYTES)};
    const uint64_t ring_d = umma_desc_sw128(sRing);          // descriptor of ring byte 0; stage s / buffer b add (s * STAGE_BYTES + b * KVB) >> 4
    auto kv_desc = [&](int stage, int buf) { return ring_d + (uint64_t)((stage * STAGE_BYTES + buf * KVB) >> 4); };
    const bool leader = elect_one();
    mbar_wait(bar_q, 0);
    mbar_wait(kv_full(0), 0);
    tc_fence_after();
    for (int w = 0; w < n_wg; ++w)
      if (leader) issue_s(tmem_base + w * TILE_COLS, qd_h[w], qd_l[w], kv_desc(0, 0), kv_desc(0, 2), n_chunks == 1 ? idesc_s_last : idesc_s, s_full(w));
    __syncwarp();
    int stage = 0, stage1 = 1 % STAGES;
    uint32_t phase1 = (1 / STAGES) & 1;
    for (int j = 0; j < n_chunks; ++j) {
      const bool more = j + 1 < n_chunks;
      const bool full = j < n_chunks - 1 || last_width == CH;           // 128-wide chunk: eight straight-line slices
      const uint64_t vh = kv_desc(stage, 1), vl = kv_desc(stage, 3), kh1 = kv_desc(stage1, 0), kl1 = kv_desc(stage1, 2);
      const uint32_t idesc1 = (j + 1 == n_chunks - 1) ? idesc_s_last : idesc_s;
      const uint32_t accum = j > 0 ? 1u : 0u;
      for (int w = 0; w < n_wg; ++w) {
        const uint32_t tS = tmem_base + w * TILE_COLS, tP = tS, tO = tS + O_COL;
        if (lane == 0 && w == 0) TRACE(2, j, 0);
        mbar_wait(p_full(w), j & 1);                     // P_w(j) is in TMEM over S_w(j), O_w rescaled if it had to be
        tc_fence_after();
        if (lane == 0) TRACE(2, j, w == 0 ? 1 : 3);
        if (leader) {
          if (full) {
#pragma unroll
            for (int k = 0; k < CH / 16; ++k) issue_o_slice(tO, tP, vh, vl, k, k > 0 ? 1u : accum);
          } else {
            for (int k = 0; k < (last_width >> 4); ++k) issue_o_slice(tO, tP, vh, vl, k, k > 0 ? 1u : accum);
          }
          tc_commit(o_done(w));
          if (!more) tc_commit(o_final(w));
        }
        __syncwarp();
        if (more) {
          mbar_wait(kv_full(stage1), phase1);            // K(j+1) has landed (returns at once for the second tile)
          tc_fence_after();
          if (leader) issue_s(tS, qd_h[w], qd_l[w], kh1, kl1, idesc1, s_full(w));
          __syncwarp();
        }
        if (lane == 0) TRACE(2, j, w == 0 ? 2 : 4);
      }
      if (leader) tc_commit(kv_empty(stage));   // K_j / V_j are free once everything issued so far has retired
      __syncwarp();
      if (++stage == STAGES) stage = 0;
      if (++stage1 == STAGES) { stage1 = 0; phase1 ^= 1; }
    }
  }
  } else {
    // ------------------------------------------------ softmax warpgroups: thread = query row
    if (X3) asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    const int w = warp >> 2;                                 // query tile of this warpgroup
    if (w < n_wg) {
      const int row = (warp & 3) * 32 + lane;               // TMEM lane == query row of this tile
      const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
      const uint32_t tS = tmem_base + w * TILE_COLS + lane_off, tO = tS + O_COL;
      const float c_log2 = 0.125f * 1.4426950408889634f;    // head_dim^-0.5 * log2(e)   (attention.py:41)
      const float tau = RESCALE_LOG2 / c_log2;
      float m_run = -INFINITY, l_run = 0.f;
      constexpr int POLY = X3 ? 0 : POLY_T;

      // ---- pass 1 over one 32-column piece: running maxima (four independent chains of 3-input maxima)
      auto piece_max = [&](auto general_tag, const uint32_t (&s)[32], const int nv, float (&mx4)[4]) {
        constexpr bool GENERAL = decltype(general_tag)::value;
        if (!GENERAL) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
#pragma unroll
            for (int u = 0; u < 4; ++u) mx4[u] = max3(mx4[u], __uint_as_float(s[i + 2 * u]), __uint_as_float(s[i + 2 * u + 1]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (i < nv) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(s[i]));
        }
      };
      // ---- pass 2 over one piece: p = 2^(s*c - m*c) for 32 scores, row-sum partials, 16-bit packing, store over the scores.
      // One-pass mode packs bf16 by TRUNCATION with one byte-permute per pair; the exponent carries +log2(1 + 2^-9), which
      // centres the truncation error (|rel| <= 2^-8, mean 0); the row sum is taken from the same values and corrected once at
      // the end.  (hi, lo) mode: FP16 pair, hi = rn(p), lo = rn(p - hi): ~22 bits of p.
      auto piece_exp = [&](auto general_tag, const int q, const uint32_t (&s)[32], const int nv, const float nmc, uint64_t (&l2)[2]) {
        constexpr bool GENERAL = decltype(general_tag)::value;
        const uint64_t c2 = pack2(c_log2, c_log2), n2 = pack2(nmc, nmc);
        float e[32];
#pragma unroll
        for (int t = 0; t < 32; t += 2) {
          const uint64_t x2 = fma2v(pack2(__uint_as_float(s[t]), __uint_as_float(s[t + 1])), c2, n2);
          if ((t & 7) < POLY) {
            // FMA-pipe 2^x: x = n + f with n = round(x) (magic-number add), |f| <= 1/2, cubic in f, n into the exponent field.
            // x is clamped to >= -125 so the exponent field stays positive (2^-125 is zero for every purpose here).
            float x0, x1;
            unpack2(x2, x0, x1);
            const uint64_t xc = pack2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
            const uint64_t magic = pack2(12582912.f, 12582912.f), nmagic = pack2(-12582912.f, -12582912.f), m1 = pack2(-1.f, -1.f);
            const uint64_t xf = add2v(xc, magic);                          // low mantissa bits = n (two's complement)
            const uint64_t fr = fma2v(add2v(xf, nmagic), m1, xc);          // f = x - n
            uint64_t pl = fma2v(pack2(0.05517164617776871f, 0.05517164617776871f), fr, pack2(0.2426111251115799f, 0.2426111251115799f));
            pl = fma2v(pl, fr, pack2(0.6932609677314758f, 0.6932609677314758f));
            pl = fma2v(pl, fr, pack2(0.9999280571937561f, 0.9999280571937561f));
            float p0, p1, f0, f1;
            unpack2(pl, p0, p1);
            unpack2(xf, f0, f1);
            e[t] = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(f0) << 23));
            e[t + 1] = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(f1) << 23));
          } else {
            float x0, x1;
            unpack2(x2, x0, x1);
            e[t] = ex2_ftz(x0);
            e[t + 1] = ex2_ftz(x1);
          }
          if (GENERAL) {
            if (t >= nv) e[t] = 0.f;
            if (t + 1 >= nv) e[t + 1] = 0.f;
          }
        }
#pragma unroll
        for (int t = 0; t < 32; t += 4) { l2[0] = add2v(l2[0], pack2(e[t], e[t + 1])); l2[1] = add2v(l2[1], pack2(e[t + 2], e[t + 3])); }
        uint32_t hi[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          if constexpr (X3) {
            const __half2 h = __floats2half2_rn(e[2 * t], e[2 * t + 1]);
            hi[t] = *reinterpret_cast<const uint32_t*>(&h);
            const float2 f = __half22float2(h);
            e[2 * t] -= f.x;
            e[2 * t + 1] -= f.y;
          } else {
            hi[t] = __byte_perm(__float_as_uint(e[2 * t]), __float_as_uint(e[2 * t + 1]), 0x7632);
          }
        }
        tc_st16(tS + 16 * q, hi);
        if constexpr (X3) {
          uint32_t lo[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            const __half2 h = __floats2half2_rn(e[2 * t], e[2 * t + 1]);
            lo[t] = *reinterpret_cast<const uint32_t*>(&h);
          }
          tc_st16(tS + P_LO_COL + 16 * q, lo);
        }
      };

      auto chunk_pipelined = [&](auto general_tag, const int j) {
        constexpr bool GENERAL = decltype(general_tag)::value;
        const int n_valid = GENERAL ? p.last_valid : CH;        // keys of this chunk that exist
        const int n_pieces = GENERAL ? (last_width + 31) >> 5 : CH / 32;
        uint32_t sa[32], sb[32];
        if (lane == 0 && (warp & 3) == 0) TRACE(w, j, 0);
        mbar_wait(s_full(w), j & 1);
        tc_fence_after();
        if (lane == 0 && (warp & 3) == 0) TRACE(w, j, 1);
        // ---- pass 1: row maximum; the load of piece q + 1 is in flight while piece q is reduced
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        tc_ld32_issue(tS, sa);
        tc_ld_wait();
        if (n_pieces > 1) tc_ld32_issue(tS + 32, sb);
        piece_max(general_tag, sa, n_valid, mx4);
        if (n_pieces > 1) {
          tc_ld_wait();
          if (n_pieces > 2) tc_ld32_issue(tS + 64, sa);
          piece_max(general_tag, sb, n_valid - 32, mx4);
          if (n_pieces > 2) {
            tc_ld_wait();
            if (n_pieces > 3) tc_ld32_issue(tS + 96, sb);
            piece_max(general_tag, sa, n_valid - 64, mx4);
            if (n_pieces > 3) {
              tc_ld_wait();
              tc_ld32_issue(tS, sa);                              // piece 0 again, for pass 2
              piece_max(general_tag, sb, n_valid - 96, mx4);
            }
          }
        }
        if (n_pieces <= 3) tc_ld32_issue(tS, sa);
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // ---- lazy rescale decision: keep the stale maximum unless the new one is > 2^8 larger
        const bool need = mx > m_run + tau;
        float alpha = 1.f;
        if (need) {
          alpha = ex2_ftz((m_run - mx) * c_log2);             // 0 on the first chunk (m_run = -inf)
          m_run = mx;
          l_run *= alpha;
        }
        const float nmc = fmaf(-m_run, c_log2, X3 ? 0.f : TRUNC_BIAS_LOG2);
        if (lane == 0 && (warp & 3) == 0) TRACE(w, j, 2);
        // ---- pass 2: exponentials; P piece q (16 columns) lands on score columns that pass 2 has already consumed
        uint64_t l2[2] = {0ull, 0ull};
        tc_ld_wait();
        if (n_pieces > 1) tc_ld32_issue(tS + 32, sb);
        piece_exp(general_tag, 0, sa, n_valid, nmc, l2);
        if (n_pieces > 1) {
          tc_ld_wait();
          if (n_pieces > 2) tc_ld32_issue(tS + 64, sa);
          piece_exp(general_tag, 1, sb, n_valid - 32, nmc, l2);
          if (n_pieces > 2) {
            tc_ld_wait();
            if (n_pieces > 3) tc_ld32_issue(tS + 96, sb);
            piece_exp(general_tag, 2, sa, n_valid - 64, nmc, l2);
            if (n_pieces > 3) {
              tc_ld_wait();
              piece_exp(general_tag, 3, sb, n_valid - 96, nmc, l2);
            }
          }
        }
        {
          float a0, a1, a2, a3;
          unpack2(l2[0], a0, a1);
          unpack2(l2[1], a2, a3);
          l_run += (a0 + a1) + (a2 + a3);
        }
        // ---- O_w rescale by the row owner (rare): P_w(j-1) V(j-1) must have retired.  S_w(j) was issued after P_w(j-1) V(j-1), so
        // o_done has completed exactly j phases by now and the parity wait is unambiguous.
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          mbar_wait(o_done(w), (j - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tc_ld32_issue(tO + 32 * h, sa);
            tc_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) sa[i] = __float_as_uint(__uint_as_float(sa[i]) * alpha);
            tc_st32(tO + 32 * h, sa);
          }
        }
        if (lane == 0 && (warp & 3) == 0) TRACE(w, j, 3);
        tc_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_local(p_full(w));
        if (lane == 0 && (warp & 3) == 0) TRACE(w, j, 4);
      };
      // One-pass mode (four CTAs per SM, 80 registers per thread): one 32-column piece in registers at a time; the TMEM round
      // trips hide behind the other three softmax warps of the scheduler.
      auto chunk_simple = [&](auto general_tag, const int j) {
        constexpr bool GENERAL = decltype(general_tag)::value;
        const int n_valid = GENERAL ? p.last_valid : CH;
        const int n_pieces = GENERAL ? (last_width + 31) >> 5 : CH / 32;
        uint32_t sa[32];
        if (lane == 0 && (warp & 3) == 0) TRACE(w, j, 0);
        mbar_wait(s_full(w), j & 1);
        tc_fence_after();
        if (lane == 0 && (warp & 3) == 0) TRACE(w, j, 1);
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int q = 0; q < CH / 32; ++q) {
          if (q < n_pieces) { tc_ld32_issue(tS + 32 * q, sa); tc_ld_wait(); piece_max(general_tag, sa, n_valid - 32 * q, mx4); }
        }
        tc_ld32_issue(tS, sa);                                  // piece 0 again, in flight across the rescale decision
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        const bool need = mx > m_run + tau;
        float alpha = 1.f;
        if (need) {
          alpha = ex2_ftz((m_run - mx) * c_log2);
          m_run = mx;
          l_run *= alpha;
        }
        const float nmc = fmaf(-m_run, c_log2, X3 ? 0.f : TRUNC_BIAS_LOG2);
        if (lane == 0 && (warp & 3) == 0) TRACE(w, j, 2);
        uint64_t l2[2] = {0ull, 0ull};
#pragma unroll
        for (int q = 0; q < CH / 32; ++q) {
          if (q < n_pieces) {
            if (q > 0) tc_ld32_issue(tS + 32 * q, sa);
            tc_ld_wait();
            piece_exp(general_tag, q, sa, n_valid - 32 * q, nmc, l2);
          }
        }
        {
          float a0, a1, a2, a3;
          unpack2(l2[0], a0, a1);
          unpack2(l2[1], a2, a3);
          l_run += (a0 + a1) + (a2 + a3);
        }
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          mbar_wait(o_done(w), (j - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            tc_ld32_issue(tO + 32 * h, sa);
            tc_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) sa[i] = __float_as_uint(__uint_as_float(sa[i]) * alpha);
            tc_st32(tO + 32 * h, sa);
          }
        }
        if (lane == 0 && (warp & 3) == 0) TRACE(w, j, 3);
        tc_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_local(p_full(w));
        if (lane == 0 && (warp & 3) == 0) TRACE(w, j, 4);
      };
      auto chunk = [&](auto general_tag, const int j) {
        if constexpr (X3) chunk_pipelined(general_tag, j);
        else chunk_simple(general_tag, j);
      };
      const bool last_general = p.last_valid != CH;
      for (int j = 0; j < n_chunks; ++j) {
        if (last_general && j == n_chunks - 1) chunk(std::true_type{}, j);
        else chunk(std::false_type{}, j);
      }
      // ---- epilogue: O / l -> 16-bit (hi[, lo]) -> staged row -> one bulk copy per plane (SPLIT: per half row).  The rows are
      // staged in this tile's own Q buffers (every S_w has retired).  o_final completes once, after the last P V (o_done's parity
      // would be ambiguous here).
      mbar_wait(o_final(w), 0);
      tc_fence_after();
      const float inv = (X3 ? 1.0f : 1.0f + 0x1p-9f) / l_run;    // one-pass mode: l was summed from values carrying the +2^-9 bias
      constexpr int NCOL = 64;
      uint8_t* const srow = base_ptr + w * TILE_BYTES + row * 128;
      uint8_t* const srow_lo = srow + NT * TILE_BYTES;
      uint32_t ro[NCOL];
      tc_ld32_issue(tO, *reinterpret_cast<uint32_t(*)[32]>(&ro[0]));
      tc_ld32_issue(tO + 32, *reinterpret_cast<uint32_t(*)[32]>(&ro[32]));
      tc_ld_wait();
#pragma unroll
      for (int g = 0; g < NCOL / 8; ++g) {
        uint32_t h2[4], l2w[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float a = __uint_as_float(ro[g * 8 + 2 * t]) * inv, c = __uint_as_float(ro[g * 8 + 2 * t + 1]) * inv;
          if constexpr (X3) {                                    // (hi, lo) FP16 pair (|O| <= max |v|: no saturation needed beyond V's own)
            const __half2 h = __floats2half2_rn(a, c);
            const float2 f = __half22float2(h);
            const __half2 l = __floats2half2_rn(a - f.x, c - f.y);
            h2[t] = *reinterpret_cast<const uint32_t*>(&h);
            l2w[t] = *reinterpret_cast<const uint32_t*>(&l);
          } else {
            const __nv_bfloat162 h = __floats2bfloat162_rn(a, c);
            h2[t] = *reinterpret_cast<const uint32_t*>(&h);
          }
        }
        *reinterpret_cast<uint4*>(srow + g * 16) = *reinterpret_cast<const uint4*>(h2);
        if (X3) *reinterpret_cast<uint4*>(srow_lo + g * 16) = *reinterpret_cast<const uint4*>(l2w);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const int q = q0 + w * 128 + row;
      if (q < p.tq_main) {
        const size_t go = ((size_t)b * T + q) * D + head * 64;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.out_hi + go), "r"(smem_u32(srow)), "n"(NCOL * 2) : "memory");
        if (X3) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.out_lo + go), "r"(smem_u32(srow_lo)), "n"(NCOL * 2) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.trace && tid == 0 && cta_lin < 8192) {
    p.trace[480 + 3 * cta_lin + 1] = globaltimer_ns();
    p.trace[480 + 3 * cta_lin + 2] = (unsigned long long)(clock64() - cta_clk0);
  }
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)V6Cfg<X3>::TMEM_COLS) : "memory");
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

// one side stream + fork / join events per device, created on first use (the library's only stream; see prv2_attention)
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; bool ok; };
SideStream g_side[64];
std::mutex g_side_mu;
SideStream* side_stream() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(g_side_mu);
  SideStream& s = g_side[dev];
  if (!s.ok) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    s.ok = true;
  }
  return &s;
}

}  // namespace

static CUresult encode_qkv_map(PFN_cuTensorMapEncodeTiled_v12000 enc, CUtensorMap* tm, const void* ptr, int B, int T, int D, int box_rows) {
  cuuint64_t dims[3] = {(cuuint64_t)3 * D, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)3 * D * 2, (cuuint64_t)T * 3 * D * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

extern "C" int prv2_attention(const prv2_bf16* qkv_hi, const prv2_bf16* qkv_lo, int B, int T, int heads, prv2_bf16* out_hi, prv2_bf16* out_lo,
                              prv2_stream_t stream) {
  PRV2_CHECK_ARG(qkv_hi && out_hi, "prv2_attention: null pointer");
  PRV2_CHECK_ARG((qkv_lo == nullptr) == (out_lo == nullptr), "prv2_attention: lo planes must both be present or absent");
  PRV2_CHECK_ARG(B > 0 && T > 0 && heads > 0 && heads <= 65535 && B <= 65535, "prv2_attention: bad shape");
  auto enc = get_encode();
  if (!enc) { set_error("prv2_attention: cuTensorMapEncodeTiled unavailable"); return PRV2_ECUDA; }
  const int D = heads * 64;
  const bool x3 = qkv_lo != nullptr;
  AttnParamsV6 p;
  memset(&p, 0, sizeof(p));
  CUresult r = encode_qkv_map(enc, &p.tm_hi, qkv_hi, B, T, D, 128);
  if (r == CUDA_SUCCESS) r = encode_qkv_map(enc, &p.tm_lo, x3 ? qkv_lo : qkv_hi, B, T, D, 128);
  const int CH = x3 ? V6Cfg<true>::CH : V6Cfg<false>::CH;
  if (r == CUDA_SUCCESS) r = encode_qkv_map(enc, &p.tmkv_hi, qkv_hi, B, T, D, CH);
  if (r == CUDA_SUCCESS) r = encode_qkv_map(enc, &p.tmkv_lo, x3 ? qkv_lo : qkv_hi, B, T, D, CH);
  if (r != CUDA_SUCCESS) { set_error("prv2_attention: cuTensorMapEncodeTiled failed (%d)", (int)r); return PRV2_ECUDA; }
  p.out_hi = (bf16*)out_hi; p.out_lo = (bf16*)out_lo;
  p.B = B; p.T = T; p.heads = heads; p.D = D;
  // key chunks: CH wide; the last one holds the remaining 1..CH keys (its MMAs run in 16-key slices)
  p.n_chunks = cdiv(T, CH);
  p.last_valid = T - CH * (p.n_chunks - 1);
  // query rows: leftover rows (<= 16) go to the SIMT tail kernel instead of a mostly empty 128-row tile
  const int rem = T % 128;
  p.tq_main = (rem != 0 && rem <= TAIL_MAX && T > 128 && T <= TAIL_MAX_T) ? T - rem : T;
  static const bool one_cta = getenv("PRV2_ATTN_ONE_CTA") != nullptr;     // diagnostics: pad the request so only one CTA fits an SM
  const int smem_bf = one_cta ? 120 * 1024 : V6Cfg<false>::SMEM, smem_x3 = V6Cfg<true>::SMEM;
  if (!g_attr_set) {
    PRV2_CUDA(cudaFuncSetAttribute(attention_v6_kernel<false, PRV2_ATTN_POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bf));
    PRV2_CUDA(cudaFuncSetAttribute(attention_v6_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_x3));
    g_attr_set = true;
  }
  static const bool want_trace = getenv("PRV2_ATTN_TRACE") != nullptr;
  static unsigned long long* d_trace = nullptr;
  if (want_trace) {
    if (!d_trace) PRV2_CUDA(cudaMalloc(&d_trace, (480 + 3 * 8192) * 8));
    PRV2_CUDA(cudaMemsetAsync(d_trace, 0, (480 + 3 * 8192) * 8, (cudaStream_t)stream));
    p.trace = d_trace;
  }
  // The leftover-row kernel (memory-bound: it re-reads K and V of every head for one query row, 31 us per 27-patch call under
  // ncu = 3 ms per frame) is independent of the tile kernel, so it runs on a side stream of the library, forked from and joined
  // back into the caller's stream with events (capturable in a CUDA graph like any fork / join); it is launched FIRST so that
  // its small CTAs find room before the tile kernel fills every SM.  PRV2_ATTN_TAIL_SERIAL=1 keeps it on the caller's stream.
  static const bool tail_serial = getenv("PRV2_ATTN_TAIL_SERIAL") != nullptr;
  const bool has_tail = p.tq_main < T;
  // (Measured, scripts/bench_attention.py: 40 -> 27 us for a one-image call and 120 -> 117 us at 12 patches, but 223 -> 233 us at
  // 27 patches, where the tail's CTAs delay the first wave of tiles: the fork is used for grids of up to ~7 waves only.)
  const long long tile_ctas = (long long)cdiv(p.tq_main, 128 * (x3 ? V6Cfg<true>::NT : V6Cfg<false>::NT)) * heads * B;
  SideStream* side = (has_tail && !tail_serial && tile_ctas <= 2048) ? side_stream() : nullptr;
  if (has_tail) {
    cudaStream_t ts = (cudaStream_t)stream;
    if (side) {
      std::lock_guard<std::mutex> lock(g_side_mu);          // record + wait as one step: callers on other host threads share the events
      PRV2_CUDA(cudaEventRecord(side->fork, (cudaStream_t)stream));
      PRV2_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
      ts = side->stream;
    }
    attention_tail_kernel<<<dim3(T - p.tq_main, heads, B), 256, 0, ts>>>((const bf16*)qkv_hi, (const bf16*)qkv_lo, (bf16*)out_hi, (bf16*)out_lo, T, heads, p.tq_main);
    PRV2_LAUNCH_CHECK();
    if (side) PRV2_CUDA(cudaEventRecord(side->join, side->stream));
  }
  const dim3 grid_bf(cdiv(p.tq_main, 128 * V6Cfg<false>::NT), heads, B);
  if (x3) attention_v6_kernel<true, 0><<<dim3(cdiv(p.tq_main, 128 * V6Cfg<true>::NT), heads, B), V6Cfg<true>::THREADS, smem_x3, (cudaStream_t)stream>>>(p);
  else attention_v6_kernel<false, PRV2_ATTN_POLY><<<grid_bf, V6Cfg<false>::THREADS, smem_bf, (cudaStream_t)stream>>>(p);
  PRV2_LAUNCH_CHECK();
  if (want_trace) {
    static unsigned long long h[3 * 32 * 5];
    PRV2_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    PRV2_CUDA(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (unsigned long long v : h) if (v && v < t0) t0 = v;
    for (int j = 0; j < p.n_chunks && j < 32; ++j) {
      fprintf(stderr, "chunk %2d |", j);
      for (int role = 0; role < 3; ++role) {
        fprintf(stderr, " %s", role == 0 ? "sm0:" : role == 1 ? "sm1:" : "mma:");
        for (int k = 0; k < 5; ++k) fprintf(stderr, " %6lld", h[(role * 32 + j) * 5 + k] ? (long long)(h[(role * 32 + j) * 5 + k] - t0) : -1ll);
        fprintf(stderr, " |");
      }
      fprintf(stderr, "\n");
    }
    // per-CTA wall time: how many CTAs, mean / max duration, kernel span, mean CTAs in flight
    static unsigned long long hc[3 * 8192];
    PRV2_CUDA(cudaMemcpy(hc, d_trace + 480, sizeof(hc), cudaMemcpyDeviceToHost));
    unsigned long long first = ~0ull, last = 0, sum = 0, mx = 0, clk = 0;
    int n = 0;
    for (int i = 0; i < 8192; ++i) {
      if (!hc[3 * i] || !hc[3 * i + 1]) continue;
      const unsigned long long d = hc[3 * i + 1] - hc[3 * i];
      sum += d; if (d > mx) mx = d; ++n; clk += hc[3 * i + 2];
      if (hc[3 * i] < first) first = hc[3 * i];
      if (hc[3 * i + 1] > last) last = hc[3 * i + 1];
    }
    if (n) {
      fprintf(stderr, "tile CTAs %d: mean %.2f us = %.0f clocks (%.2f GHz), max %.2f us, kernel span %.2f us, mean CTAs in flight %.1f\n", n, sum / 1e3 / n,
              (double)clk / n, (double)clk / (double)sum, mx / 1e3, (last - first) / 1e3, (double)sum / (double)(last - first));
      for (int i = 0; i < n && i < 4000; i += n / 12 + 1)
        fprintf(stderr, "  cta %4d start %8.2f us dur %6.2f us (%llu clocks)\n", i, (hc[3 * i] - first) / 1e3, (hc[3 * i + 1] - hc[3 * i]) / 1e3, hc[3 * i + 2]);
    }
  }
  if (side) PRV2_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, side->join, 0));
  return PRV2_OK;
}
