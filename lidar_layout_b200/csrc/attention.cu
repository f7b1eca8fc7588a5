// Fused flash-style attention for head dim 32 on tcgen05 / TMEM / TMA (sm_100a).
//
// Replaces QKVAttentionLegacy.forward (reference lidm/modules/diffusion/openaimodel.py:358-374) and the core of
// CrossAttention.forward (reference lidm/modules/attention.py:170-193): per head
//   w = softmax_fp32((q s)^T (k s)),  a = w v,   s = ch^-1/4  (folded into the packed q / k weight rows),
// without materialising the (B*heads, T, T) score tensor.  q / k / v are read straight out of the packed (B, T, 3C)
// output of the qkv GEMM (or, for cross-attention, q from its own GEMM and k / v from the context projection); V is
// consumed as an MN-major tcgen05 operand, so nothing is transposed.
//   v5: persistent kernel, two 128-row query tiles per work item (one when T is an odd multiple of 128), P and O
//       resident in tensor memory - see namespace v5.
//   xs: CUDA-core kernel for cross-attention against a handful of context tokens.
#include <cstdlib>

#include <type_traits>

#include "common.h"
#include "ptx.cuh"

namespace lidm {

namespace {

constexpr int D = 32;
constexpr int BQ = 128;
constexpr int Q_BYTES = BQ * D * 2;        // 8 KiB, rows of 64 B (SWIZZLE_64B)

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Blackwell's packed fp32 pair instructions (FFMA2 / FADD2): one issue slot for two lanes of work, which is what the
// softmax warps run short of (per score: scale+subtract, exponential, row sum, bf16 pack, row max).
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// 2^x for a pair on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial, max relative error 7.6e-5, far
// below the bf16 rounding of P): takes a share of the exponentials off the MUFU pipe, which bounds this kernel (head
// dim 32: one exponential per 128 tensor-core flops).  5 issue slots per element against MUFU's 8 busy clocks.
__device__ __forceinline__ void ex2_poly2(float& y0, float& y1, float x0, float x1) {
  constexpr float MAGIC = 12582912.f;         // 1.5 * 2^23: the integer part of x lands in the low mantissa bits
  x0 = fmaxf(x0, -125.f);
  x1 = fmaxf(x1, -125.f);
  float t0, t1, u0, u1, f0, f1, p0, p1;
  add2(t0, t1, x0, x1, MAGIC, MAGIC);
  add2(u0, u1, t0, t1, -MAGIC, -MAGIC);
  add2(f0, f1, x0, x1, -u0, -u1);             // f in [-0.5, 0.5]
  fma2(p0, p1, f0, f1, 0.05520551f, 0.05520551f, 0.24261397f, 0.24261397f);
  fma2(p0, p1, p0, p1, f0, f1, 0.69325477f, 0.69325477f);
  fma2(p0, p1, p0, p1, f0, f1, 0.9999277f, 0.9999277f);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// One 32-score chunk of a row: p = 2^(s*log2e - mb) as 16 bf16 pairs into pk, row sums into (s0..s3).  POLYP of every
// 8 pairs (evenly spread) take the polynomial instead of MUFU.EX2.
// TRUNC (bf16 only): P is cut to bf16 by taking the upper halves of the two fp32 values (one PRMT, one issue slot) instead of
// round-to-nearest packing (F2FP: two).  The numerator P*V and the row sum (ones column of the same MMA) are formed from the
// SAME truncated P, so the downward bias of truncation cancels in O / l; the rounding noise keeps its variance.
#ifndef LIDM_ATTN_TRUNC
#define LIDM_ATTN_TRUNC 1
#endif
// sc: what turns a score into a base-2 exponent (log2 e, or 1 when the q / k weights carry it); FAST: sc == 1 and the row's
// reference is 0, so the score IS the exponent - no scale-and-subtract at all.
template <int POLYP, bool SUM, bool F16, bool TRUNC = false, bool FAST = false>
__device__ __forceinline__ void exp_chunk(const uint32_t (&sv)[32], float sc, float mb, uint32_t* pk, float& s0, float& s1,
                                          float& s2, float& s3) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float x0, x1, p0, p1;
    if (FAST) { x0 = __uint_as_float(sv[2 * i]); x1 = __uint_as_float(sv[2 * i + 1]); }
    else fma2(x0, x1, __uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1]), sc, sc, -mb, -mb);
    if (((i & 7) * POLYP) % 8 < POLYP) {
      ex2_poly2(p0, p1, x0, x1);
    } else {
      p0 = ex2(x0);
      p1 = ex2(x1);
    }
    if (SUM) {
      if (i & 1) add2(s2, s3, s2, s3, p0, p1);
      else add2(s0, s1, s0, s1, p0, p1);
    }
    if (TRUNC && !F16) pk[i] = __byte_perm(__float_as_uint(p0), __float_as_uint(p1), 0x7632);
    else pk[i] = pack_h<F16>(p0, p1);
  }
}

// =====================================================================================================
// v5 (every self-attention, and cross-attention against more than 16 context tokens): persistent kernel, one CTA per SM walking (query block, head,
// sample) work items; NG 128-row query tiles per item (one softmax warpgroup each) share every K/V tile of BKV_ keys.
//   * O accumulates in TMEM across an item's whole K/V loop (tcgen05.mma accumulate); the running max is only
//     refreshed, and O rescaled in place (tcgen05.ld / scale / tcgen05.st), when some row's max grew by more than 2^8 -
//     the result is exact because the row sum l is kept against the same (stale) max.  Nothing in the softmax loop
//     waits on the P*V product.
//   * P never touches shared memory: the softmax threads write it (bf16 pairs) into tensor memory with tcgen05.st and
//     the P*V product takes its A operand from there (tcgen05.mma [d], [a_tmem], b_desc).  With P in shared memory the
//     N=32 P*V instructions are bound by streaming the 32 KiB P tile through the shared-memory port (measured 45 clk
//     per M128 N32 K16 instruction against 16 ideal, tests/microbench/umma_rate.cu) on top of the 32 KiB the softmax
//     warps store.
//   * S needs a single TMEM buffer per group because each softmax thread drains its whole S row into registers first
//     and hands the buffer straight back.
//   * one MMA-issuing thread per group (tcgen05.mma issue blocks until the tensor pipe accepts the instruction, so one
//     thread serving both groups delays one group's S / P*V behind the other's), fixed issue order S(n+1), P(n)V(n),
//     parked mbarrier waits.
//   * TMEM, the barriers and the K/V ring live across items (cumulative phases) and the next item's queries load into
//     a second buffer, so its loads and first S = Q K^T overlap the current item's tail (an item per CTA costs ~3.6 us
//     of un-overlapped prologue/epilogue: T = 2048, B = 64: 738 -> 715 us).
//   * the two softmax warps that share a scheduler either take strict turns at the exponential phase (PP: mbarrier
//     ping-pong, short items: T = 512 145 -> 128 us) or free-run, started a fraction of a tile apart at every item
//     (long items: T = 2048 715 us against 794 us with the ping-pong).
//   * a quarter of the exponentials run as a packed polynomial on the FMA pipe (exp_chunk).
// TMEM: S_g at columns g*BKV_, P_g (BKV_/2 columns) at NG*BKV_ + g*BKV_/2, O_g at NG*BKV_*3/2 + g*32.
namespace v5 {

constexpr int KV_ST = 4;
constexpr float RESCALE_LOG2 = 8.f;

// HS (half split): every 128-row query tile is served by TWO softmax warpgroups, each owning 64 of the tile's 128 key
// columns (the row max is exchanged through shared memory), i.e. four softmax warps per scheduler instead of two - the
// exponential phase is bound by dependency stalls that two warps per scheduler cannot cover (ncu: top stall "wait").
template <int NG, int BKV_, bool HS = false>
struct Cfg {
  static constexpr int KB = BKV_ * 64;                        // bytes of one K (or V) tile
  static constexpr int OFF_Q = 0;                             // two buffers: the next work item's queries load early
  static constexpr int OFF_K = OFF_Q + 2 * NG * Q_BYTES;
  static constexpr int OFF_V = OFF_K + KV_ST * KB;
  static constexpr int OFF_ONES = OFF_V + KV_ST * KB;          // constant second N-atom of the P*V B operand (row sums)
  static constexpr int OFF_XMAX = OFF_ONES + KB;                // HS: [parity][group][half][128 rows] half-row maxima
  static constexpr int OFF_BAR = OFF_XMAX + (HS ? 2 * NG * 2 * 128 * 4 : 0);
  static constexpr int SMEM_TOTAL = OFF_BAR + 512 + 1024;
  static constexpr int SPLIT = HS ? 2 : 1;                      // softmax warpgroups per query tile
  static constexpr uint32_t P_COL = NG * BKV_;
  static constexpr uint32_t O_COL = P_COL + NG * (BKV_ / 2);
  static constexpr uint32_t O_STRIDE = 64;                    // 32 columns of O, column 32 = row sum (N = 48 P*V), padding
  static constexpr uint32_t TMEM_COLS = (O_COL + NG * O_STRIDE) <= 256 ? 256 : 512;
  static constexpr int THREADS = 128 + NG * 128 * SPLIT;
  static constexpr int NCH = BKV_ / 32 / SPLIT;               // 32-column chunks of an S row per softmax thread
  static_assert(O_COL + NG * O_STRIDE <= 512, "TMEM budget");
  static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");
};

template <int NG, int BKV_, int POLYP, bool PP, bool F16, bool HS>
__global__ void __launch_bounds__((Cfg<NG, BKV_, HS>::THREADS), 1)
attention_d32_v5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                        bf16* __restrict__ out, int out_ld, int T, int q_col, int k_col, int v_col, int kv_len,
                        int n_qblk, int heads, int n_items, int submax, float sscale) {
  using L = Cfg<NG, BKV_, HS>;
  static_assert(!HS || (!PP && BKV_ == 128), "half split: free-running groups, 128-key tiles");
  // Row sums on the tensor pipe: the P*V product runs with N = 48, the B operand's second 32-column atom being a constant
  // tile whose first column is all ones, so column 32 of O accumulates sum_k P[row][k] (of the bf16-rounded P the
  // numerator uses) for free - an N = 48 instruction costs what N = 32 does - and the softmax warps drop one FADD2
  // per element pair from their dispatch-bound exponential phase.
  constexpr bool MMA_ROWSUM = true;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);   // [2]
  uint64_t* q_empty = q_full + 2;          // [2]: every group's last S of the item has been issued and completed
  uint64_t* kv_full = q_empty + 2;         // [KV_ST]
  uint64_t* kv_empty = kv_full + KV_ST;    // [KV_ST]
  uint64_t* s_ready = kv_empty + KV_ST;    // [NG]
  uint64_t* s_free = s_ready + NG;         // [NG]
  uint64_t* p_ready = s_free + NG;         // [NG]
  uint64_t* pv_done = p_ready + NG;        // [NG]
  uint64_t* stagger = pv_done + NG;        // [NG]: group g-1 -> group g, once per CTA
  uint64_t* xu_go = stagger + NG;          // [2 groups][4 schedulers]: exp-phase ping-pong between the two warps of a scheduler
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xu_go + 8);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform as far as the compiler can tell
  const int lane = threadIdx.x & 31;
  const int nkv = (kv_len + BKV_ - 1) / BKV_;   // rows past kv_len are zero-filled by TMA and masked to -inf below
  // persistent CTA: work item = (query block, head, sample); this CTA takes items blockIdx.x, blockIdx.x + gridDim.x, ...
  // TMEM, the barriers and the K/V ring live across items (all phases are counted cumulatively), so the next item's
  // Q / K / V loads and its first S = Q K^T overlap the tail of the current one
  const int n_my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto item_coords = [&](int it, int& q0, int& head, int& b) {
    const int item = it * gridDim.x + blockIdx.x;
    const int qb = item % n_qblk;
    const int hb = item / n_qblk;
    q0 = qb * (NG * BQ);
    head = hb % heads;
    b = hb / heads;
  };

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmKV);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], NG); }
    for (int s = 0; s < KV_ST; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], NG); }
    for (int g = 0; g < NG; ++g) { mbar_init(&s_ready[g], 1); mbar_init(&s_free[g], 4 * L::SPLIT); mbar_init(&stagger[g], 4 * L::SPLIT); }
    for (int i = 0; i < NG; ++i) { mbar_init(&p_ready[i], 4 * L::SPLIT); mbar_init(&pv_done[i], 1); }
    for (int i = 0; i < 8; ++i) mbar_init(&xu_go[i], 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, L::TMEM_COLS); tmem_relinquish(); }
  if (MMA_ROWSUM) {
    // ones atom: [BKV_ keys][32 columns] bf16, 64-byte rows, SWIZZLE_64B (16-byte chunk c of row r sits at chunk
    // c ^ ((r >> 1) & 3)); column 0 = 1.0, the rest 0
    uint4* ones = reinterpret_cast<uint4*>(smem + L::OFF_ONES);
    for (int i = threadIdx.x; i < BKV_ * 4; i += blockDim.x) {
      const int r = i >> 2, cphys = i & 3;
      const int c = cphys ^ ((r >> 1) & 3);
      ones[i] = make_uint4(c == 0 ? (F16 ? 0x00003c00u : 0x00003f80u) : 0u, 0u, 0u, 0u);   // 1.0 in the low half = element 0
    }
    fence_proxy_async();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp < 4) {
    if (HS) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    else if (NG == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // the whole warp walks the loop (warp-uniform operands stay in uniform registers); one elected lane issues
      {
        const bool leader = elect_one() != 0;
        auto load_q = [&](int it) {
          int q0, head, b;
          item_coords(it, q0, head, b);
          uint64_t* bar = &q_full[it & 1];
          if (leader) {
            mbar_arrive_expect_tx(bar, NG * Q_BYTES);
#pragma unroll
            for (int i = 0; i < NG; i += 2)                                     // 256 query rows per box
              tma_load_3d(smem + L::OFF_Q + ((it & 1) * NG + i) * Q_BYTES, &tmQ, bar, q_col + head * D, q0 + i * BQ, b);
          }
          __syncwarp();
        };
        if (n_my > 0) load_q(0);
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < n_my; ++it) {
          if (it + 1 < n_my) {
            if (it + 1 >= 2) mbar_wait(&q_empty[(it + 1) & 1], ((it - 1) >> 1) & 1);   // buffer last used by item it-1
            load_q(it + 1);
          }
          int q0, head, b;
          item_coords(it, q0, head, b);
          for (int j = 0; j < nkv; ++j) {
            mbar_wait(&kv_empty[s], ph ^ 1);
            if (leader) {
              mbar_arrive_expect_tx(&kv_full[s], 2 * L::KB);
              tma_load_3d(smem + L::OFF_K + s * L::KB, &tmKV, &kv_full[s], k_col + head * D, j * BKV_, b);
              tma_load_3d(smem + L::OFF_V + s * L::KB, &tmKV, &kv_full[s], v_col + head * D, j * BKV_, b);
            }
            __syncwarp();
            if (++s == KV_ST) { s = 0; ph ^= 1; }
          }
        }
      }
    } else if (warp - 1 < NG) {
      // one MMA-issuing thread per group (tcgen05.mma issue blocks until the tensor pipe accepts the instruction, so a
      // single thread serving both groups delays one group's S / P*V behind the other's): each polls only its own
      // barriers; a K/V stage goes back to the TMA warp when every group has committed its P*V on it
      {
        const bool leader = elect_one() != 0;
        const int g = warp - 1;
        constexpr uint32_t idesc_s = make_idesc_h<F16>(BQ, BKV_);
        constexpr uint32_t idesc_o = make_idesc_h<F16>(BQ, MMA_ROWSUM ? 48 : D) | (1u << 16);   // V is an MN-major B operand
        // tiles are counted cumulatively over all of this CTA's items (barrier phases never reset); the issue order is
        // fixed - S(n+1), then P(n) V(n) - and every wait parks the thread (mbar_wait), so the issuer takes no issue
        // slots from the softmax warps that share its scheduler
        const int total = n_my * nkv;
        const uint32_t dS = tmem_base + g * BKV_;
        const uint32_t tP = tmem_base + L::P_COL + g * (BKV_ / 2);
        const uint32_t dO = tmem_base + L::O_COL + g * L::O_STRIDE;
        int it_s = 0, j_s = 0;                 // item / tile-in-item of the next S to issue
        auto issue_s = [&](int ns) {
          const int st = ns % KV_ST;
          mbar_wait(&kv_full[st], (ns / KV_ST) & 1);
          if (j_s == 0) mbar_wait(&q_full[it_s & 1], (it_s >> 1) & 1);
          if (ns > 0) mbar_wait(&s_free[g], (ns - 1) & 1);
          tcgen05_fence_after();
          const uint64_t qdesc = make_kmajor_desc<64>(smem_u32(smem + L::OFF_Q + ((it_s & 1) * NG + g) * Q_BYTES));
          const uint64_t kdesc = make_kmajor_desc<64>(smem_u32(smem + L::OFF_K + st * L::KB));
          if (leader) {
            umma_bf16_ss(dS, qdesc, kdesc, idesc_s, 0);
            umma_bf16_ss(dS, qdesc + 2, kdesc + 2, idesc_s, 1);
            umma_commit(&s_ready[g]);
            if (j_s == nkv - 1) umma_commit(&q_empty[it_s & 1]);   // this group no longer reads the item's queries
          }
          __syncwarp();
          if (++j_s == nkv) { j_s = 0; ++it_s; }
        };
        if (total > 0) issue_s(0);
        for (int np = 0, j_p = 0; np < total; ++np) {
          if (np + 1 < total) issue_s(np + 1);
          const int st = np % KV_ST;
          // p_ready of an item's first tile also implies that the softmax warps have read the previous item's O
          mbar_wait(&p_ready[g], np & 1);
          tcgen05_fence_after();
          // MN-major B operand: 32-column atoms along N at the leading-byte-offset stride; the second atom is the
          // constant ones tile, wherever this stage's V tile sits
          uint64_t vdesc = make_kmajor_desc<64>(smem_u32(smem + L::OFF_V + st * L::KB));
          if (MMA_ROWSUM) {
            const uint64_t lbo = (uint64_t)(L::OFF_ONES - (L::OFF_V + st * L::KB)) >> 4;
            vdesc = (vdesc & ~(0x3FFFull << 16)) | (lbo << 16);
          }
          if (leader) {
#pragma unroll
            for (int kk = 0; kk < BKV_ / 16; ++kk) {
              const uint64_t vb = vdesc + (uint64_t)((kk * 1024) >> 4);
              umma_bf16_ts(dO, tP + kk * 8, vb, idesc_o, (j_p > 0 || kk != 0) ? 1u : 0u);   // 16 bf16 = 8 columns
            }
            umma_commit(&pv_done[g]);
            umma_commit(&kv_empty[st]);      // this group is done with the K/V tile
          }
          __syncwarp();
          if (++j_p == nkv) j_p = 0;
        }
      }
    }
  } else {
    if (HS) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    else if (NG == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int g = (warp - 4) / (4 * L::SPLIT);     // softmax group = query tile
    const int half = HS ? ((warp - 4) >> 2) & 1 : 0;   // which 64 of the tile's 128 key columns this warpgroup owns
    const int qd = warp & 3;                       // TMEM lane quadrant
    const int row = qd * 32 + lane;                // row inside the 128-row tile
    const int col0 = half * (BKV_ / 2);            // first key column of this thread inside a K/V tile
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + g * BKV_ + col0;
    const uint32_t tO = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + L::O_COL + g * L::O_STRIDE;
    const uint32_t tP = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + L::P_COL + g * (BKV_ / 2) + col0 / 2;
    float* xmax = reinterpret_cast<float*>(smem + L::OFF_XMAX);
    // sscale turns scores into base-2 exponents: log2 e, or 1 when the packed q / k weights already carry it (the engine's
    // own weights do).  Then, in the streamed bf16 mode, a row whose first-tile maximum lies within 2^+-24 takes reference 0
    // and its scores go to the exponential as they are (fast tiles: one FFMA2 per pair less in the dispatch-bound phase).
    // Streamed tiles move the reference one tile late, when a tile's maximum has left the one the reference was set from by
    // more than 2^24.  A reference that is not 0 sits 2^40 ABOVE that maximum (P = 2^(s - m) <= 2^-40 while nothing has grown):
    // bf16 P and the fp32 row sums / accumulators reach 2^127, so a row maximum may grow by 2^130 (90 nats) from one 128-key
    // tile to the next before anything overflows, while scores more than 2^86 (60 nats) below the maximum flush to zero -
    // they carry less than 2^-86 of the row sum (tests/test_gpu_ops.py::test_attention_reference_regimes walks the regimes).
    const float LOG2E = sscale;
    constexpr float REF_RANGE = 24.f;
    const bool fast_ok = !HS && !F16 && (submax & 2) != 0 && (submax & 8) == 0 && sscale == 1.f;
    const float rescale_th = (!HS && !F16 && (submax & 2) != 0) ? REF_RANGE : RESCALE_LOG2;
    const float ref_bias = (!HS && !F16 && (submax & 2) != 0) ? 40.f / sscale : 0.f;   // in score units
    int n = 0;                                     // cumulative tile index over this CTA's items (barrier phases)
    const bool restagger = nkv >= 8;
    constexpr bool PINGPONG = PP;
    // An item's epilogue (read O and the row sums out of TMEM, normalise, store) is deferred into the NEXT item's first
    // tile, after that tile's exponentials: the last P*V of the item completes while the softmax warps already work on
    // the next item's S(0) (issued early thanks to the second Q buffer), instead of being waited for with nothing to do.
    // p_ready of that first tile is only signalled after the read, so the P*V that overwrites O cannot overtake it.
    int pq0 = 0, phead = 0, pb = 0;
    bool pending = false;
    float l = 0.f;
    auto item_epilogue = [&]() {
      if (HS && half != 0) return;                   // the partner warpgroup stores the tile
      uint32_t o[32];
      tmem_ld_32x32b_x32(tO, o);
      if (MMA_ROWSUM) l = __uint_as_float(tmem_ld_32x32b_x1(tO + 32));
      tmem_ld_wait();
      tcgen05_fence_before();
      const float inv = 1.f / l;
      bf16* op = out + ((size_t)pb * T + pq0 + g * 128 + row) * out_ld + phead * D;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_h<F16>(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
        u.y = pack_h<F16>(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
        u.z = pack_h<F16>(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
        u.w = pack_h<F16>(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
        reinterpret_cast<uint4*>(op)[i] = u;
      }
    };
    for (int it = 0; it < n_my; ++it) {
    int q0, head, b;
    item_coords(it, q0, head, b);
    float m = 0.f, mx = 0.f, rprev = 0.f;         // reference subtracted from the scores; the maximum it was set from; last tile's maximum
    float lcur = 0.f;                               // running row sum of this item when it is not taken from the MMA
    for (int j = 0; j < nkv; ++j, ++n) {
      mbar_wait(&s_ready[g], n & 1);
      tcgen05_fence_after();
      uint32_t sv[L::NCH][32];
      // Streamed tiles (bf16 P, every tile after an item's first, no ragged tail): the exponentials are taken against the
      // reference m carried over from the earlier tiles, so nothing in the tile waits for its own maximum - the first
      // 32-column chunk is loaded alone, the other loads fly while its exponentials run, and this tile's (sub-sampled)
      // maximum is gathered on the side for the NEXT tile's rescale decision.  A reference that lags one tile is as good as
      // any other (P and the row sum share it; bf16 / fp32 keep their precision ~85 nats above it) - the per-tile chain
      // "all loads -> maximum -> exponentials" was a third of a softmax warp's time with the MUFU pipe idle.
      bool streamed = false;
      if constexpr (!HS && !F16) streamed = (submax & 2) != 0 && j > 0 && kv_len - j * BKV_ >= BKV_;
      if (streamed) {
        if (__any_sync(0xffffffffu, (rprev - mx) * LOG2E > rescale_th)) {
          // rare: the previous tile raised the row maximum past the threshold - refresh m and rescale O in TMEM
          mbar_wait(&pv_done[g], (n - 1) & 1);   // every P*V issued so far (the previous tile's included) has completed
          tcgen05_fence_after();
          mx = fmaxf(mx, rprev);
          const float mn = mx + ref_bias;
          const float alpha = ex2((m - mn) * LOG2E);
          uint32_t o[32];
          tmem_ld_32x32b_x32(tO, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x32(tO, o);
          if (MMA_ROWSUM) {
            const uint32_t lr = tmem_ld_32x32b_x1(tO + 32);
            tmem_ld_wait();
            tmem_st_32x32b_x1(tO + 32, __float_as_uint(__uint_as_float(lr) * alpha));
          }
          tmem_st_wait();
          lcur *= alpha;
          m = mn;
        }
        tmem_ld_32x32b_x32(tS, sv[0]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 1; c < L::NCH; ++c) tmem_ld_32x32b_x32(tS + c * 32, sv[c]);
      } else {
#pragma unroll
      for (int c = 0; c < L::NCH; ++c) tmem_ld_32x32b_x32(tS + c * 32, sv[c]);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[g]);      // S_g(j+1) may overwrite the TMEM buffer now
      if (kv_len - j * BKV_ < BKV_) {               // ragged last tile (cross-attention context): mask the padding keys
        const int valid = kv_len - j * BKV_ - col0;
#pragma unroll
        for (int c = 0; c < L::NCH; ++c) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) sv[c][i] = 0xff800000u;   // -inf
        }
      }
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
      if ((submax & 1) && !F16 && !HS) {
        // Sub-sampled reference: the running value m only has to stay within the exponent range of the true row maximum
        // (P = 2^(s - m) and the row sum l are formed against the SAME m, so any m gives the same quotient; in bf16 / fp32 a
        // score 2^100 above m still neither overflows nor loses precision), so the maximum over every fourth column pair
        // (16 of the 64 FMNMX3 of a tile, a tenth of the softmax warps' issue slots) is as good a reference as the exact one.
        // IEEE-half P (F16) keeps the exact maximum: its exponent range is too short for values far below the reference.
#pragma unroll
        for (int c = 0; c < L::NCH; ++c) {
#pragma unroll
          for (int i = 0; i < 4; ++i) m0 = max3(m0, __uint_as_float(sv[c][8 * i + 0]), __uint_as_float(sv[c][8 * i + 1]));
        }
      } else {
#pragma unroll
        for (int c = 0; c < L::NCH; ++c) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            m0 = max3(m0, __uint_as_float(sv[c][8 * i + 0]), __uint_as_float(sv[c][8 * i + 1]));
            m1 = max3(m1, __uint_as_float(sv[c][8 * i + 2]), __uint_as_float(sv[c][8 * i + 3]));
            m2 = max3(m2, __uint_as_float(sv[c][8 * i + 4]), __uint_as_float(sv[c][8 * i + 5]));
            m3 = max3(m3, __uint_as_float(sv[c][8 * i + 6]), __uint_as_float(sv[c][8 * i + 7]));
          }
        }
      }
      float r = max3(fmaxf(m0, m1), m2, m3);
      if (HS) {
        // the row max over all 128 keys: swap half-row maxima with the partner warp (same lane quadrant, other half);
        // parity double-buffered, one 64-thread named barrier per (group, quadrant)
        float* xm = xmax + (((n & 1) * NG + g) * 2) * 128;
        xm[half * 128 + row] = r;
        named_bar_sync(1 + g * 4 + qd, 64);
        r = fmaxf(r, xm[(half ^ 1) * 128 + row]);
      }
      rprev = r;
      if (j == 0) {
        const bool zero_ref = fast_ok && fabsf(r) <= REF_RANGE;      // (fast_ok: scores are base-2 exponents)
        mx = zero_ref ? 0.f : r;
        m = zero_ref ? 0.f : r + ref_bias;
      } else if (__any_sync(0xffffffffu, (r - mx) * LOG2E > rescale_th)) {
        // rare: refresh the running max of every row of this warp and rescale O in TMEM
        mbar_wait(&pv_done[g], (n - 1) & 1);   // every P*V issued so far has completed
        tcgen05_fence_after();
        mx = fmaxf(mx, r);
        const float mn = mx + ref_bias;
        const float alpha = ex2((m - mn) * LOG2E);
        if (!HS || half == 0) {       // both halves take the same decision (same rows, same maxima); one rescales O
          uint32_t o[32];
          tmem_ld_32x32b_x32(tO, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x32(tO, o);
          if (MMA_ROWSUM) {
            const uint32_t lr = tmem_ld_32x32b_x1(tO + 32);
            tmem_ld_wait();
            tmem_st_32x32b_x1(tO + 32, __float_as_uint(__uint_as_float(lr) * alpha));
          }
          tmem_st_wait();
        }
        lcur *= alpha;
        m = mn;
      }
      // start the groups a fraction of a tile apart (and, on long items, re-establish the offset at every item: the
      // phase relation is only neutrally stable and drifts back towards lock-step otherwise)
      // exp-phase ping-pong: warp 4+q (group 0) and warp 8+q (group 1) share scheduler q and its MUFU unit; they take
      // strict turns at the exponential phase (A(0), B(0), A(1), B(1), ...), so each runs it at the full MUFU rate
      // while the other drains TMEM / takes the row max / waits for S, instead of both halving each other's rate
      if (PINGPONG) {
        if (g == 0) { if (n > 0) mbar_wait(&xu_go[qd], (n - 1) & 1); }
        else mbar_wait(&xu_go[4 + qd], n & 1);
      } else if (j == 0 && g > 0 && nkv > 1 && (it == 0 || restagger)) {
        mbar_wait(&stagger[g], restagger ? (it & 1) : 0);
      }
      }   // !streamed
      const float mb = m * LOG2E;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      uint32_t pkk[32];
      float rs = -INFINITY;                          // streamed tile: its sub-sampled maximum, for the next tile
      const bool fast = fast_ok && __all_sync(0xffffffffu, m == 0.f);
      auto chunks = [&](auto fast_tag) {
        constexpr bool FAST = decltype(fast_tag)::value;
#pragma unroll
      for (int c = 0; c < L::NCH; ++c) {
        uint32_t* pk = &pkk[HS ? 0 : (c & 1) * 16];
        if (streamed) {
#pragma unroll
          for (int i = 0; i < 4; ++i) rs = max3(rs, __uint_as_float(sv[c][8 * i + 0]), __uint_as_float(sv[c][8 * i + 1]));
        }
        exp_chunk<POLYP, !MMA_ROWSUM, F16, LIDM_ATTN_TRUNC != 0, FAST>(sv[c], LOG2E, mb, pk, s0, s1, s2, s3);
        if (streamed && c == 0) {                    // the other chunks have landed behind these exponentials
          tmem_ld_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[g]);
        }
        if (HS) {
          // half split: 16 columns (32 keys) per store keeps the packed pairs out of the register budget of 96
          if (c == 0 && n > 0) {
            mbar_wait(&pv_done[g], (n - 1) & 1);
            tcgen05_fence_after();
          }
          tmem_st_32x32b_x16(tP + c * 16, pk);
        } else if (c & 1) {
          if (c == 1 && n > 0) {
            // the previous P*V of this group (the previous tile's, or the previous item's last) must have drained P
            // before it is overwritten; by now half of this tile's exponentials are done, so the wait is normally free
            mbar_wait(&pv_done[g], (n - 1) & 1);
            tcgen05_fence_after();
          }
          tmem_st_32x32b_x32(tP + (c >> 1) * 32, pkk);   // 64 keys = 32 columns of bf16 pairs
        }
        if (!PINGPONG && c == (HS ? 0 : L::NCH / NG - 1 + (L::NCH / NG == 0)) && j == 0 && (it == 0 || restagger) && g + 1 < NG && lane == 0)
          mbar_arrive(&stagger[g + 1]);
      }
      };
      if (fast) chunks(std::true_type{});
      else chunks(std::false_type{});
      if (streamed) rprev = rs;
      if (PINGPONG && lane == 0) mbar_arrive(&xu_go[(g == 0 ? 4 : 0) + qd]);   // hand the MUFU unit to the partner warp
      if (!MMA_ROWSUM) lcur += (s0 + s1) + (s2 + s3);
      tmem_st_wait();
      if (j == 0 && pending) {                      // previous item: its last P*V was waited for at c == 1 above
        item_epilogue();
        pending = false;
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[g]);
    }
    pq0 = q0; phead = head; pb = b;
    if (!MMA_ROWSUM) l = lcur;
    pending = true;
    }   // work items
    if (pending) {
      mbar_wait(&pv_done[g], (n - 1) & 1);   // the last item's last P*V (n already counts it)
      tcgen05_fence_after();
      item_epilogue();
    }
  }
  __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem_base, L::TMEM_COLS); }
}

// q: (B, T, q_ld) rows with the heads at columns q_col + head*32; k / v: (B, kv_rows, kv_ld) rows at columns k_col / v_col.
template <int NG, int BKV_, int POLYP, bool PP, bool F16, bool HS = false>
void launch_f(const bf16* q, int q_ld, int q_col, const bf16* kv, int kv_ld, int k_col, int v_col, int kv_rows, const View& out,
            int B, int T, int heads, cudaStream_t s, float sscale) {
  using L = Cfg<NG, BKV_, HS>;
  static bool configured = false;
  if (!configured) {
    LIDM_CUDA_CHECK(cudaFuncSetAttribute(attention_d32_v5_kernel<NG, BKV_, POLYP, PP, F16, HS>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM_TOTAL));
    configured = true;
  }
  LIDM_REQUIRE(T % (NG * BQ) == 0 && kv_rows >= 1, "attention tile shape");
  CUtensorMap tmQ = make_tma_3d(q, q_ld, T, B, (uint64_t)q_ld * 2, (uint64_t)T * q_ld * 2, D, NG >= 2 ? 256 : 128, 64);
  CUtensorMap tmKV = make_tma_3d(kv, kv_ld, kv_rows, B, (uint64_t)kv_ld * 2, (uint64_t)kv_rows * kv_ld * 2, D, BKV_, 64);
  const int n_qblk = T / (NG * BQ);
  const int n_items = n_qblk * heads * B;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    LIDM_CUDA_CHECK(cudaGetDevice(&dev));
    LIDM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = n_items < num_sms ? n_items : num_sms;     // one persistent CTA per SM
  // bit 0: sub-sampled maximum (T = 2048: 698 -> 640 us); bit 1: streamed tiles against the lagging reference (same box:
  // 689 -> 628 us; requesting the next tile's first chunk during the last chunk's exponentials on top of it: 656-668 us,
  // not kept).  A/B switch; 0 = exact maximum of every tile before its exponentials.
  static const int submax = getenv("LIDM_ATTN_SUBMAX") ? atoi(getenv("LIDM_ATTN_SUBMAX")) : 3;
  launch_pdl(attention_d32_v5_kernel<NG, BKV_, POLYP, PP, F16, HS>, dim3(grid), dim3(L::THREADS), L::SMEM_TOTAL, s, tmQ, tmKV, out.p,
             out.ld, T, q_col, k_col, v_col, kv_rows, n_qblk, heads, n_items, submax, sscale);
  LIDM_COUNT_LAUNCH(1);
}

template <int NG, int BKV_, int POLYP, bool PP = (NG == 2), bool HS = false>
void launch(const bf16* q, int q_ld, int q_col, const bf16* kv, int kv_ld, int k_col, int v_col, int kv_rows, const View& out,
            int B, int T, int heads, cudaStream_t s, float sscale) {
  if (out.f16) launch_f<NG, BKV_, POLYP, PP, true, HS>(q, q_ld, q_col, kv, kv_ld, k_col, v_col, kv_rows, out, B, T, heads, s, sscale);
  else launch_f<NG, BKV_, POLYP, PP, false, HS>(q, q_ld, q_col, kv, kv_ld, k_col, v_col, kv_rows, out, B, T, heads, s, sscale);
}

}  // namespace v5

// =====================================================================================================
// Cross-attention against a handful of context tokens (cam2lidar: L = 4 camera tokens): the tensor-core kernel would
// spend a whole 128-key tile, its barriers and 124 masked exponentials per row on L dot products.  Here one thread owns
// one (pixel pair, head): q . k_l, softmax over l and sum_l p_l v_l in fp32 registers, K / V of the sample staged once
// per CTA in shared memory as floats (rows padded to 33 per head: the heads of a warp hit different banks, the pixels
// of a warp broadcast).  Memory-bound: reads q, writes the output, nothing else.
namespace xs {

constexpr int LMAX = 16;
constexpr int HP = 33;   // padded floats per (token, head) row

__global__ void __launch_bounds__(256)
xattn_small_kernel(const bf16* __restrict__ q, int q_ld, const bf16* __restrict__ kv, int kv_ld, int k_col, int v_col, int L,
                   bf16* __restrict__ out, int out_ld, int T, int heads, float sscale) {
  extern __shared__ float skv[];                     // K: [L][heads][HP], then V: same
  const int b = blockIdx.y;
  const int C = heads * D;
  float* sk = skv;
  float* sv = skv + L * heads * HP;
  const int C8 = C >> 3;
  for (int i = threadIdx.x; i < 2 * L * C8; i += blockDim.x) {     // 16-byte loads, all of a thread's in flight at once
    const int isv = i >= L * C8;
    const int r = isv ? i - L * C8 : i;
    const int l = r / C8, c = (r - l * C8) * 8;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(kv + ((size_t)b * L + l) * kv_ld + (isv ? v_col : k_col) + c));
    float* dst = (isv ? sv : sk) + (l * heads + c / D) * HP + (c % D);
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16(uu[k]);
      dst[2 * k] = f.x;
      dst[2 * k + 1] = f.y;
    }
  }
  __syncthreads();
  const float LOG2E = sscale;                        // log2 e, or 1 when the q weights carry it
  const int pairs = (T / 2) * heads;                 // work item = (pixel pair, head); heads fastest -> coalesced rows
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < pairs; idx += gridDim.x * blockDim.x) {
    const int hd = idx % heads, t0 = (idx / heads) * 2;
    float qf[2][D];
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
      const uint4* qp = reinterpret_cast<const uint4*>(q + ((size_t)b * T + t0 + pp) * q_ld + hd * D);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 u = __ldg(qp + i);
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16(uu[k]);
          qf[pp][i * 8 + 2 * k] = f.x;
          qf[pp][i * 8 + 2 * k + 1] = f.y;
        }
      }
    }
    float sc[2][LMAX];
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
      if (l < L) {
        const float* kr = sk + (l * heads + hd) * HP;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float kd = kr[d];
          a0 = fmaf(qf[0][d], kd, a0);
          a1 = fmaf(qf[1][d], kd, a1);
        }
        sc[0][l] = a0; sc[1][l] = a1;
        mx[0] = fmaxf(mx[0], a0); mx[1] = fmaxf(mx[1], a1);
      }
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
      if (l < L) {
        sc[0][l] = ex2((sc[0][l] - mx[0]) * LOG2E); sum[0] += sc[0][l];
        sc[1][l] = ex2((sc[1][l] - mx[1]) * LOG2E); sum[1] += sc[1][l];
      }
    }
    float* of0 = qf[0];                              // q is dead: reuse its registers for the output rows
    float* of1 = qf[1];
#pragma unroll
    for (int d = 0; d < D; ++d) { of0[d] = 0.f; of1[d] = 0.f; }
#pragma unroll
    for (int l = 0; l < LMAX; ++l) {
      if (l < L) {
        const float* vr = sv + (l * heads + hd) * HP;
        const float p0 = sc[0][l], p1 = sc[1][l];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const float vd = vr[d];
          of0[d] = fmaf(p0, vd, of0[d]);
          of1[d] = fmaf(p1, vd, of1[d]);
        }
      }
    }
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
      const float inv = 1.f / sum[pp];
      const float* o = pp ? of1 : of0;
      uint4* op = reinterpret_cast<uint4*>(out + ((size_t)b * T + t0 + pp) * out_ld + hd * D);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16(o[8 * i + 0] * inv, o[8 * i + 1] * inv);
        u.y = pack_bf16(o[8 * i + 2] * inv, o[8 * i + 3] * inv);
        u.z = pack_bf16(o[8 * i + 4] * inv, o[8 * i + 5] * inv);
        u.w = pack_bf16(o[8 * i + 6] * inv, o[8 * i + 7] * inv);
        op[i] = u;
      }
    }
  }
}

}  // namespace xs

}  // namespace

void launch_attention_d32_packed(const bf16* qkv, const View& out, int B, int T, int heads, cudaStream_t s, bool log2_scores) {
  const int C = heads * D;
  const float sscale = log2_scores ? 1.f : 1.4426950408889634f;   // q . k is already a base-2 exponent when the weights carry log2 e
  LIDM_REQUIRE(T % 128 == 0, "attention: T must be a multiple of 128");
  LIDM_REQUIRE(out.hl == 0 && out.hr == 0 && out.H * out.W == T && out.B == B && out.C == C, "attention out view");
  LIDM_REQUIRE(out.ld % 8 == 0, "attention out ld");
  // A quarter of the exponentials (2 of every 8 pairs) run as a polynomial on the FMA pipe.  Measured on B200
  // (tests/microbench/pipe_rate.cu, exp_loop2.cu): the scheduler dispatches one warp instruction per clock and FFMA2 /
  // FADD2 / F2FP / FMNMX3 hold it for two, so a polynomial pair costs 16 dispatch clocks - as many as the two MUFU.EX2
  // it replaces keep the MUFU pipe busy - and the optimum is where both run out together: 2 of 8 (exp phase 2130 ->
  // 1820 clk per tile pair in isolation; T = 2048 784 -> 738 us, T = 512 121 -> 119 us in the U-Net at B = 64);
  // 3 of 8 is already slower (also with the row sums on the tensor pipe: T = 2048 703 us at 2 of 8, 711 us at 3 of 8).
  // LIDM_ATTN_POLY=0 turns it off (A/B runs).
  static const int poly = getenv("LIDM_ATTN_POLY") ? atoi(getenv("LIDM_ATTN_POLY")) : 3;
#define LIDM_ATTN_ARGS qkv, 3 * C, 0, qkv, 3 * C, C, 2 * C, T, out, B, T, heads, s, sscale
  // Free-running (staggered) softmax groups at every length: with the packed softmax and parked waits the strict
  // ping-pong no longer pays even for short items (T = 512, same box: 125.0 us free-running, 128.5 us ping-pong); it
  // stays available as the PP template flag and serves the single-tile cross-attention items.  Odd multiples of 128
  // (T = 128: the 4x32 level) take one query tile per item; the persistent kernel beats a one-tile-per-CTA kernel with
  // two CTAs per SM there too (35.4 against 39.5 us at B = 64).
  static const int hs = getenv("LIDM_ATTN_HS") ? atoi(getenv("LIDM_ATTN_HS")) : 0;
  if (T % 256 == 0) {
    if (hs) v5::launch<2, 128, 2, false, true>(LIDM_ATTN_ARGS);
    else if (poly == 3) v5::launch<2, 128, 3, false>(LIDM_ATTN_ARGS);
    else if (poly == 4) v5::launch<2, 128, 4, false>(LIDM_ATTN_ARGS);
    else if (poly == 1) v5::launch<2, 128, 1, false>(LIDM_ATTN_ARGS);
    else if (poly) v5::launch<2, 128, 2, false>(LIDM_ATTN_ARGS);
    else v5::launch<2, 128, 0, false>(LIDM_ATTN_ARGS);
  } else {
    if (poly) v5::launch<1, 128, 2, false>(LIDM_ATTN_ARGS);
    else v5::launch<1, 128, 0, false>(LIDM_ATTN_ARGS);
  }
#undef LIDM_ATTN_ARGS
}

// CrossAttention.forward core (reference lidm/modules/attention.py:170-193) for head dim 32: q (B,T,q_ld) against a
// short context k/v (B, L, kv_ld), L not necessarily a multiple of 128 (TMA zero-fills, the kernel masks).
void launch_cross_attention_d32(const bf16* q, int q_ld, const bf16* kv, int kv_ld, int k_col, int v_col, int L,
                                const View& out, int B, int T, int heads, cudaStream_t s, bool log2_scores) {
  LIDM_REQUIRE(T % 128 == 0 && L >= 1, "cross attention: T must be a multiple of 128");
  const float sscale = log2_scores ? 1.f : 1.4426950408889634f;
  static const bool no_small = getenv("LIDM_XATTN_TC") != nullptr;   // A/B switch: always take the tensor-core kernel
  if (L <= xs::LMAX && !no_small && !out.f16 && q_ld % 8 == 0 && out.ld % 8 == 0 && kv_ld % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0 &&
      (size_t)2 * L * heads * xs::HP * sizeof(float) <= 48 * 1024) {
    LIDM_REQUIRE(out.hl == 0 && out.hr == 0 && out.H * out.W == T && out.B == B && out.C == heads * D, "attention out view");
    const int pairs = (T / 2) * heads;
    int gx = (pairs + 255) / 256;
    const int cap = (148 * 8 + B - 1) / B;           // ~4 waves at 2 CTAs per SM: measured faster than one long-running wave
    if (gx > cap) gx = cap;
    const size_t smem = (size_t)2 * L * heads * xs::HP * sizeof(float);
    xs::xattn_small_kernel<<<dim3(gx, B), 256, smem, s>>>(q, q_ld, kv, kv_ld, k_col, v_col, L, out.p, out.ld, T, heads, sscale);
    LIDM_CUDA_CHECK(cudaGetLastError());
    LIDM_COUNT_LAUNCH(1);
    return;
  }
  if (T % 256 == 0) v5::launch<2, 128, 0>(q, q_ld, 0, kv, kv_ld, k_col, v_col, L, out, B, T, heads, s, sscale);
  else v5::launch<1, 128, 0>(q, q_ld, 0, kv, kv_ld, k_col, v_col, L, out, B, T, heads, s, sscale);
}

}  // namespace lidm
