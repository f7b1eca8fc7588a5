// Implicit-GEMM circular convolution on tcgen05 / TMEM fed by TMA (sm_100a).
//
// Replaces, for the LiDM sampling path, CircularConv2d.forward (reference lidm/modules/basic.py:52-59; every
// Conv2d of UNetModel, openaimodel.py, and of Decoder, model_lidm.py) and the 1x1 Conv1d/Conv2d/nn.Linear-like
// GEMMs around attention.  GEMM view:  M = B*H*W output pixels (128-pixel tiles = Wbox x Hbox patches of one
// sample), N = Cout, K = taps * Cin walked tap by tap in 64-channel slices.
//   A tile : one TMA box (64 ch, Wbox, Hbox, 1) of the channels-last bf16 activation, shifted by the tap; the
//            circular W wrap is a materialised halo column, the zero H padding is TMA out-of-bounds fill.
//   B tile : one TMA box (64 k, BN rows) of the packed weights [Cout][tap][Cin].
//   D      : fp32 accumulator in TMEM (BN columns x 128 lanes), read back with tcgen05.ld by 4 epilogue warps.
// Warp roles: warp0 = TMA producer, warp1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue.
// Epilogue fusions: bias, per-sample timestep-embedding add, residual add, halo write, transposed V^T store for
// attention, fp32 NCHW/NHWC stores, and the DDIM update (ddim.py:191-206) for the U-Net's final conv.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.h"
#include "ddim_math.cuh"
#include "ptx.cuh"

namespace lidm {

namespace {

#ifndef LIDM_EPI16
#define LIDM_EPI16 0
#endif
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB

// developer timeline (compile with -DLIDM_GEMM_TRACE_ON=1, run with LIDM_GEMM_TRACE=path): CTA 0 of the resident-weight
// kernel writes clock64 stamps [row][tile][8] - how the per-tile latency chains documented below were measured
#ifndef LIDM_GEMM_TRACE_ON
#define LIDM_GEMM_TRACE_ON 0
#endif
#if LIDM_GEMM_TRACE_ON
#define LIDM_TRACE(rowi, t, k)                                                                                        \
  do {                                                                                                                  \
    if (RESK > 0 && p.trace != nullptr && blockIdx.x == 0 && (t) < 64) p.trace[((size_t)(rowi) * 64 + (t)) * 8 + (k)] = clock64(); \
  } while (0)
#else
#define LIDM_TRACE(rowi, t, k) do { } while (0)
#endif

struct GemmKernelParams {
  int tiles_w, tiles_per_img, Wbox, Hbox;
  int bbox;                  // samples per 128-row tile (> 1 when one sample has fewer than 128 pixels), B = batch size
  int B;
  int hl;
  int sx, sy;                // convolution stride (input pixel = output pixel * stride + tap offset)
  int ntaps;                 // virtual taps = spatial taps x operand-split segments (<= 27)
  int8_t dx[27], dy[27];
  int a_coff[27];            // channel offset of the A plane read by this virtual tap (hi / lo plane)
  int b_koff[27];            // first K column of this virtual tap in the B matrix
  int kchunks;
  int k2chunks, hl2, b2_koff;  // second A operand (GemmEpilogue::a2): K chunks, its left halo, its first B column
  int a2_diag;               // the a2 block of B is the identity: a channel tile reads only its own BN / 64 chunks of a2
  int N, H, W;
  int wt_batched;
  const float* bias;
  const float* rowadd;
  int rowadd_ld;
  const bf16* res;
  float res_scale;           // out = acc + bias + res_scale * residual (R2DM: (x + f(x)) / sqrt 2 with f's scale folded into W)
  int n_fast;                // tile order: output-channel tiles of one pixel tile back to back (A read from HBM once)
  int res_ld, res_hl, res_Wp;
  const float* res_f32;      // fp32 channels-last residual (precise mode), no halo
  int res_f32_ld;
  int out_f32_ld;            // channel stride of out_f32_nhwc (concat views)
  bf16* out;
  int out_ld, out_hl, out_hr, out_Wp;
  int geglu;                 // GEGLU epilogue: columns come in blocks of [16 values | 16 gates], out = value * gelu(gate)
  float* gst;                // GroupNorm granule statistics of the output (View::gst), or null
  int gst_ld, gst_slots, gst_slot0;
  int split_n;
  long long* trace;          // developer timeline (LIDM_GEMM_TRACE): [cta][tile][8] clock64 stamps, resident-weight kernel only
  bf16* out_t;
  float* out_f32_nchw;
  float* out_f32_nhwc;
  const float* ddim_x;
  const float* ddim_noise;
  float* ddim_x_prev;
  float* ddim_pred_x0;
  const float* ddim_coef;
};

// Exact-form GELU 0.5 g (1 + erf(g / sqrt 2)) with erf(x) = tanh(x (a + b x^2 + c x^4)) (minimax fit, |erf error| < 3.7e-5,
// |GELU error| < 5.5e-5 absolute - far below the bf16 rounding of the output) on ONE MUFU op (tanh.approx) + 6 FMA-pipe
// ops: the GEGLU epilogue runs it 64 times per thread and tile and is what bounds those GEMMs (the previous
// Abramowitz-Stegun form took two MUFU ops + 14 others).  x^2 is clamped so the fitted polynomial never leaves the
// range where it is monotone (|g| > 6: erf = +-1 to 1e-9).
__device__ __forceinline__ float gelu_erf_fast(float g) {
  const float g2 = fminf(g * g, 36.f);
  const float u = g * fmaf(g2, fmaf(g2, -0.00031580628f, 0.036798257f), 0.7977178f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float h = 0.5f * g;
  return fmaf(h, t, h);
}

// =====================================================================================================
// Persistent variant: one CTA per SM walks tiles (output-channel tiles of one pixel tile back to back, so concurrently
// running CTAs share the A tile in L2; the weights are L2-resident anyway); two TMEM accumulator buffers let the MMA warp start tile i+1 while the epilogue warps drain tile i; the
// tile's bias + timestep-embedding row is staged once in shared memory; bf16 outputs leave through a swizzled
// shared-memory staging buffer and TMA stores (full 128-byte lines), double-buffered across tiles.
// RESK > 0 (resident weights): a 1x1 GEMM whose whole K = RESK weight slice of one output-channel tile fits in shared
// memory keeps it there for the life of the CTA - every CTA owns ONE output-channel tile and walks pixel tiles, the
// ring of STAGES buffers then carries A chunks only.  With streamed weights a short-K tile re-reads its 2 * BN * K bytes of
// B for every 128 pixels (128 KB of B against 64 KB of A at K = 256, BN = 256) and the 3-stage ring bounds the tile time
// by the load latency (ncu: the epilogue warps spend 59 % of their samples waiting for the accumulator).
// PAIR: two CTAs of a cluster work as one tcgen05 cta_group::2 unit on two pixel tiles of the SAME channel tile: each CTA
// loads its own A rows and HALF of the B rows, the leader's MMAs (M = 256) read both halves, each CTA's tensor memory gets
// its own 128 rows.  A 128x256 tile re-reads 32 KB of weights per 16 KB of activations per K chunk and the 3x3 convs are
// bound by exactly that L2 -> shared-memory traffic (profiles/r02f_summary.md); the pair halves the weight share.
template <int BN, int STAGES, int RESK = 0, bool PAIR = false>
struct PersistLayout {
  static constexpr int B_STAGE_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;
  static constexpr int B_STRIDE = (B_STAGE_BYTES + 1023) / 1024 * 1024;
  static constexpr int B_BUFS = RESK > 0 ? RESK / BK : STAGES;       // resident: one buffer per K chunk
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = STAGES * A_STAGE_BYTES;
  static constexpr int OUT_BOXES = BN >= 64 ? BN / 64 : 0;          // 64-channel TMA store boxes per tile
  static constexpr int OUT_BUF_BYTES = (RESK > 0 ? 1 : OUT_BOXES) * BM * 128;   // 2-byte staging: one tile (resident: one 64-column box per group)
  static constexpr int OUT_BUFS = (BN > 128 || RESK > 256) ? 1 : 2; // staging buffers (double-buffered when they fit)
  static constexpr int OUT_OFF = B_OFF + B_BUFS * B_STRIDE;
  static constexpr int BIAS_OFF = OUT_OFF + OUT_BUFS * OUT_BUF_BYTES;
  static constexpr int BAR_OFF = BIAS_OFF + 2 * (BN < 32 ? 32 : BN) * 4;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
  static constexpr uint32_t ACC_COLS = BN;                          // TMEM columns per accumulator buffer
  static constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  // epilogue warps: 4 cover the 128 TMEM lanes; wide tiles use two such groups, each draining half of the columns
  // (BN = 256: sixteen - four per scheduler, 64 columns per thread: the short-K 1x1 GEMMs are bound by the epilogue's
  // dependent chain TMEM load -> bias/residual -> pack -> staging store, which two warps per scheduler do not cover)
  static constexpr int EPI_WARPS = BN >= 256 ? LIDM_EPI16 * 8 + 8 : (BN >= 64 ? 8 : 4);
  static constexpr int EPI_THREADS = EPI_WARPS * 32;
  static constexpr int THREADS = 64 + EPI_THREADS;
  static constexpr int COLS = BN / (EPI_WARPS / 4);                 // columns drained by one epilogue thread
};

template <int BN, int STAGES, bool F16, int RESK = 0, bool PAIR = false>
__global__ void __launch_bounds__((PersistLayout<BN, STAGES, RESK, PAIR>::THREADS), 1)
conv_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmA2,
                         const GemmKernelParams p, int num_m_tiles,
                         int num_tiles, int use_tma_store) {
  using L = PersistLayout<BN, STAGES, RESK, PAIR>;
  static_assert(!(PAIR && RESK > 0), "CTA pairs: streamed weights only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator drained
  uint64_t* bres_bar = tempty_bar + 2;        // resident weights have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres_bar + 1);
  float* sbias = reinterpret_cast<float*>(smem + L::BIAS_OFF);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform as far as the compiler can tell
  const int lane = threadIdx.x & 31;
  const int num_it = p.ntaps * p.kchunks + (p.a2_diag ? BN / BK : p.k2chunks);
  const int num_n_tiles = num_tiles / num_m_tiles;
  // tile walk: streamed weights - tiles blockIdx.x, + gridDim.x, ... of the (pixel tile, channel tile) grid; resident weights -
  // the CTA owns channel tile blockIdx.x % num_n_tiles and walks pixel tiles (the launcher makes gridDim.x a multiple of it)
  // CTA pairs: the walk is over PAIRS of pixel tiles (2 mp + rank), blockIdx.x / 2 = the cluster
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const int num_m_units = PAIR ? num_m_tiles / 2 : num_m_tiles;
  const int tile_first = RESK > 0 ? (int)blockIdx.x / num_n_tiles : (PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x);
  const int tile_step = RESK > 0 ? (int)gridDim.x / num_n_tiles : (PAIR ? (int)gridDim.x / 2 : (int)gridDim.x);
  const int tile_end = RESK > 0 ? num_m_tiles : num_m_units * num_n_tiles;
  const int res_n0 = ((int)blockIdx.x % num_n_tiles) * BN;
  auto tile_coords = [&](int tile, int& m_tile, int& n0) {
    if (RESK > 0) { m_tile = tile; n0 = res_n0; return; }
    const int mu = p.n_fast ? tile / num_n_tiles : tile % num_m_units;
    n0 = (p.n_fast ? tile % num_n_tiles : tile / num_m_units) * BN;
    m_tile = PAIR ? 2 * mu + (int)cta_rank : mu;
  };

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (use_tma_store) prefetch_tensormap(&tmO);
    if (p.k2chunks) prefetch_tensormap(&tmA2);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], RESK > 0 ? 4 : (PAIR ? 2 * L::EPI_WARPS : L::EPI_WARPS));   // pair: both CTAs' epilogues release the leader
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_2sm(tmem_slot, L::TMEM_COLS); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, L::TMEM_COLS); tmem_relinquish(); }
  }
  tcgen05_fence_before();
  if (PAIR) cluster_sync_all();      // barrier inits and the allocation are visible to the peer before any remote arrive / MMA
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                  // everything below reads / writes tensors of earlier kernels
  pdl_launch_dependents();     // the next kernel may run its prologue while this one drains

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp walks the loop (warp-uniform control flow and operands: the compiler keeps tile coordinates, stage
    // indices and descriptors in uniform registers); one elected lane issues the asynchronous instructions.  With the loop
    // under `if (lane == 0)` every UTMALDG / UTCHMMA / UTCBAR was wrapped in an ELECT / R2UR.BROADCAST / BRA.U.ANY
    // retry loop over the "divergent" operands, several hundred clocks of single-thread overhead per K chunk.
    {
      const bool leader = elect_one() != 0;
      int s = 0; uint32_t ph = 0;
      if (RESK > 0 && tile_first < tile_end) {
        if (leader) {
          mbar_arrive_expect_tx(bres_bar, L::B_BUFS * L::B_STAGE_BYTES);
#pragma unroll
          for (int kc = 0; kc < L::B_BUFS; ++kc)
            tma_load_3d(smem + L::B_OFF + kc * L::B_STRIDE, &tmB, bres_bar, kc * BK, res_n0, 0);
        }
        __syncwarp();
      }
      if constexpr (RESK > 0) {
        // resident weights: the ring is two whole-tile buffers of RESK / 64 chunks with ONE barrier pair per tile - the
        // issuing threads' per-chunk scalar work (~300 clocks of dependent uniform-datapath instructions per chunk, measured)
        // would otherwise outlast the 256 clocks the tensor pipe needs for a 128-wide chunk
        constexpr int KCH = RESK / BK;
        static_assert(STAGES == 2 * KCH, "resident weights: two tile buffers");
        int lt = 0;
        for (int tile = tile_first; tile < tile_end; tile += tile_step, ++lt) {
          const int tb = lt & 1;
          const int bt = tile / p.tiles_per_img;
          const int r = tile - bt * p.tiles_per_img;
          const int th = r / p.tiles_w;
          const int h0 = th * p.Hbox, w0 = (r - th * p.tiles_w) * p.Wbox;
          mbar_wait(&empty_bar[tb], ((lt >> 1) & 1) ^ 1);
          if (leader) {
            LIDM_TRACE(0, lt, 0);
            mbar_arrive_expect_tx(&full_bar[tb], KCH * A_STAGE_BYTES);
#pragma unroll
            for (int kc = 0; kc < KCH; ++kc)
              tma_load_4d(smem + L::A_OFF + (tb * KCH + kc) * A_STAGE_BYTES, &tmA, &full_bar[tb], p.a_coff[0] + kc * BK, w0 + p.hl,
                          h0, bt);
            LIDM_TRACE(0, lt, 1);
          }
          __syncwarp();
        }
      } else
      for (int tile = tile_first; tile < tile_end; tile += tile_step) {
        int m_tile, n0;
        tile_coords(tile, m_tile, n0);
        const int bn0 = n0 + (PAIR ? (int)cta_rank * (BN / 2) : 0);       // this CTA's rows of the B tile
        const int bt = m_tile / p.tiles_per_img;
        const int b = bt * p.bbox;
        const int r = m_tile - bt * p.tiles_per_img;
        const int th = r / p.tiles_w;
        const int h0 = th * p.Hbox, w0 = (r - th * p.tiles_w) * p.Wbox;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int x = w0 * p.sx + p.hl + p.dx[tap], y = h0 * p.sy + p.dy[tap];
          const int acoff = p.a_coff[tap], bkoff = p.b_koff[tap];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (leader) {
              if (RESK > 0) {
                mbar_arrive_expect_tx(&full_bar[s], A_STAGE_BYTES);
                tma_load_4d(smem + L::A_OFF + s * A_STAGE_BYTES, &tmA, &full_bar[s], acoff + kc * BK, x, y, b);
              } else if (PAIR) {
                // both CTAs' bytes are counted on the LEADER's barrier (its MMA thread waits for the whole pair)
                const uint32_t fb = mapa_shared(smem_u32(&full_bar[s]), 0);
                if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (A_STAGE_BYTES + L::B_STAGE_BYTES));
                tma_load_4d_2sm(smem + L::A_OFF + s * A_STAGE_BYTES, &tmA, fb, acoff + kc * BK, x, y, b);
                tma_load_3d_2sm(smem + L::B_OFF + s * L::B_STRIDE, &tmB, fb, bkoff + kc * BK, bn0, p.wt_batched ? b : 0);
              } else {
                mbar_arrive_expect_tx(&full_bar[s], A_STAGE_BYTES + L::B_STAGE_BYTES);
                tma_load_4d(smem + L::A_OFF + s * A_STAGE_BYTES, &tmA, &full_bar[s], acoff + kc * BK, x, y, b);
                tma_load_3d(smem + L::B_OFF + s * L::B_STRIDE, &tmB, &full_bar[s], bkoff + kc * BK, n0,
                            p.wt_batched ? b : 0);
              }
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
        }
        const int k2lo = p.a2_diag ? n0 / BK : 0, k2hi = p.a2_diag ? (n0 + BN) / BK : p.k2chunks;
        for (int kc = k2lo; kc < k2hi; ++kc) {        // second A operand: one unshifted tap after the main K range
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (leader && PAIR) {
            const uint32_t fb = mapa_shared(smem_u32(&full_bar[s]), 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (A_STAGE_BYTES + L::B_STAGE_BYTES));
            tma_load_4d_2sm(smem + L::A_OFF + s * A_STAGE_BYTES, &tmA2, fb, kc * BK, w0 + p.hl2, h0, b);
            tma_load_3d_2sm(smem + L::B_OFF + s * L::B_STRIDE, &tmB, fb, p.b2_koff + kc * BK, bn0, p.wt_batched ? b : 0);
          } else if (leader) {
            mbar_arrive_expect_tx(&full_bar[s], A_STAGE_BYTES + L::B_STAGE_BYTES);
            tma_load_4d(smem + L::A_OFF + s * A_STAGE_BYTES, &tmA2, &full_bar[s], kc * BK, w0 + p.hl2, h0, b);
            tma_load_3d(smem + L::B_OFF + s * L::B_STRIDE, &tmB, &full_bar[s], p.b2_koff + kc * BK, n0,
                        p.wt_batched ? b : 0);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp in the loop, one elected lane issues)
    {
      const bool leader = elect_one() != 0;
      constexpr uint32_t idesc = make_idesc_h<F16>(BM, BN);
      constexpr uint32_t idesc_pair = make_idesc_h<F16>(2 * BM, BN);
      int s = 0; uint32_t ph = 0;
      int lt = 0;
      if (RESK > 0 && tile_first < tile_end) mbar_wait(bres_bar, 0);
      if constexpr (RESK > 0) {
        constexpr int KCH = RESK / BK;
        for (int tile = tile_first; tile < tile_end; tile += tile_step, ++lt) {
          const int ab = lt & 1;                             // accumulator and tile buffer
          if (leader) LIDM_TRACE(0, lt, 2);
          mbar_wait(&tempty_bar[ab], ((lt >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
          if (leader) LIDM_TRACE(0, lt, 3);
          mbar_wait(&full_bar[ab], (lt >> 1) & 1);           // the whole A tile has landed
          tcgen05_fence_after();
          if (leader) LIDM_TRACE(0, lt, 4);
          const uint32_t d = tmem_base + ab * L::ACC_COLS;
          const uint64_t adesc0 = make_kmajor_desc<128>(smem_u32(smem + L::A_OFF + ab * KCH * A_STAGE_BYTES));
          const uint64_t bdesc0 = make_kmajor_desc<128>(smem_u32(smem + L::B_OFF));
          if (leader) {
#pragma unroll
            for (int kc = 0; kc < KCH; ++kc) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_bf16_ss(d, adesc0 + (uint64_t)((kc * A_STAGE_BYTES) >> 4) + 2 * k, bdesc0 + (uint64_t)((kc * L::B_STRIDE) >> 4) + 2 * k,
                             idesc, (kc | k) != 0);
            }
            umma_commit(&empty_bar[ab]);
            umma_commit(&tfull_bar[ab]);
            LIDM_TRACE(0, lt, 6);
          }
          __syncwarp();
        }
      } else if (!PAIR || cta_rank == 0)       // pair: the leader issues for both CTAs
      for (int tile = tile_first; tile < tile_end; tile += tile_step, ++lt) {
        const int ab = lt & 1;
        mbar_wait(&tempty_bar[ab], ((lt >> 1) & 1) ^ 1);   // epilogue(s) have drained this accumulator
        tcgen05_fence_after();
        const uint32_t d = tmem_base + ab * L::ACC_COLS;
        for (int it = 0; it < num_it; ++it) {
          mbar_wait(&full_bar[s], ph);
          tcgen05_fence_after();
          const uint64_t adesc = make_kmajor_desc<128>(smem_u32(smem + L::A_OFF + s * A_STAGE_BYTES));
          const uint64_t bdesc = make_kmajor_desc<128>(smem_u32(smem + L::B_OFF + s * L::B_STRIDE));
          if (leader) {
            if (PAIR) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_bf16_ss_2sm(d, adesc + 2 * k, bdesc + 2 * k, idesc_pair, (it | k) != 0);
              umma_commit_2sm(&empty_bar[s], 3);           // frees the stage in both CTAs
            } else {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_bf16_ss(d, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
              umma_commit(&empty_bar[s]);
            }
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (leader) {
          if (PAIR) umma_commit_2sm(&tfull_bar[ab], 3);
          else umma_commit(&tfull_bar[ab]);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;              // TMEM lane quadrant this warp may access
    const int wg = (warp - 2) >> 2;      // column group (0 or 1)
    const int col_lo = wg * L::COLS;     // this thread drains tile columns [col_lo, col_lo + COLS)
    const int row = q * 32 + lane;
    const int e = threadIdx.x - 64;      // 0..EPI_THREADS-1
    const int rps = BM / p.bbox;         // tile rows per sample
    const int rb = row / rps;            // sample of this row inside the tile (0 unless bbox > 1)
    const int hh = (row - rb * rps) / p.Wbox;
    const int ww = (row - rb * rps) - hh * p.Wbox;
    const int HW = p.H * p.W;
    float coef[5] = {0, 0, 0, 0, 0};
    if (p.ddim_x_prev != nullptr) {
#pragma unroll
      for (int i = 0; i < 5; ++i) coef[i] = __ldg(p.ddim_coef + i);
    }
    constexpr int CH = BN < 32 ? BN : 32;
    constexpr int SB = BN < 32 ? 32 : BN;
    int lt = 0;
    if constexpr (RESK > 0) {
      // ---- resident-weight tiles (BN = 128, 2-byte output through TMA stores, no residual).  The per-tile epilogue is a
      // serial chain of synchronisations (accumulator barrier, TMEM load, proxy fence, staging-buffer reuse, named barrier,
      // store issue: ~1500 clocks of latency measured with clock64 stamps against ~800 of data work for 64 columns per
      // thread), so TWO groups of four warps take alternate tiles and run two such chains at once: group g owns
      // accumulator g and ONE 64-column staging box (16 KB - the shared memory goes to the A ring, whose depth bounds the
      // tile rate: STAGES chunks per ~2000-clock load latency), drains its tile in two 64-column passes, each staged and
      // stored on its own; the group's first thread issues the stores.  The channel tile never changes, so the bias is
      // staged once; pixel coordinates advance incrementally; staging and bias traffic uses shared-space accesses.
      static_assert(BN == 128 && L::EPI_WARPS == 8 && L::OUT_BUFS == 2, "resident-weight epilogue: 128-column tiles, two groups of four warps");
      const int grp = wg;                    // tile parity this warp serves
      const int eg = e & 127;                // thread index inside the group
      for (int i = e; i < BN; i += L::EPI_THREADS) sbias[i] = (res_n0 + i < p.N && p.bias != nullptr) ? __ldg(p.bias + res_n0 + i) : 0.f;
      named_bar_sync(1, L::EPI_THREADS);
      const int n0 = res_n0;
      const uint32_t sbias_s = smem_u32(sbias);
      uint8_t* stage_out = smem + L::OUT_OFF + grp * L::OUT_BUF_BYTES;
      const uint32_t stage_s = smem_u32(stage_out);
      const uint32_t taddr_row = tmem_base + grp * L::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
      const uint32_t row_s = stage_s + row * 128;
      // this group's tiles: tile_first + grp * tile_step, then every 2 * tile_step; (sample, tile-in-sample) kept incrementally
      const int gstep = 2 * tile_step;
      const int step_b = gstep / p.tiles_per_img, step_r = gstep - step_b * p.tiles_per_img;
      int tile = tile_first + grp * tile_step;
      int b0 = tile / p.tiles_per_img, r = tile - b0 * p.tiles_per_img;
      const bool geglu = p.geglu != 0;       // GEGLU: a pass yields 32 output channels, the box leaves after the second pass
      bool buf_busy = false;                 // a store from this group's box may still be reading it
      for (int gl = 0; tile < tile_end; tile += gstep, ++gl) {
        const int th = p.tiles_w == 1 ? r : r / p.tiles_w;
        const int h0 = th * p.Hbox, w0 = (r - th * p.tiles_w) * p.Wbox;
        const int b = b0;
        if (eg == 0) LIDM_TRACE(1 + grp, gl, 0);
        mbar_wait(&tfull_bar[grp], gl & 1);
        tcgen05_fence_after();
        if (eg == 0) LIDM_TRACE(1 + grp, gl, 1);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t raw[2][32];
#pragma unroll
          for (int c = 0; c < 2; ++c) tmem_ld_32x32b_x32(taddr_row + hf * 64 + c * 32, raw[c]);
          tmem_ld_wait();
          if (hf == 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[grp]);
            if (eg == 0) LIDM_TRACE(1 + grp, gl, 3);
          }
          if (buf_busy && (!geglu || hf == 0)) {
            if (eg == 0) tma_store_wait_read<0>();     // the previous store of this group has left the box
            named_bar_sync(4 + grp, 128);
          }
          if (n0 + hf * 64 < p.N) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int cc = hf * 64 + c * 32;
              float v[32];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t = ld_shared_f4(sbias_s + (cc + 4 * j) * 4);
                v[4 * j] = __uint_as_float(raw[c][4 * j]) + t.x;
                v[4 * j + 1] = __uint_as_float(raw[c][4 * j + 1]) + t.y;
                v[4 * j + 2] = __uint_as_float(raw[c][4 * j + 2]) + t.z;
                v[4 * j + 3] = __uint_as_float(raw[c][4 * j + 3]) + t.w;
              }
              if (geglu) {
                const int gbase = (cc >> 1) >> 3;      // output channel cc / 2 of the tile's 64: 16-byte chunk index
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  float o[8];
#pragma unroll
                  for (int k = 0; k < 8; ++k) o[k] = v[i * 8 + k] * gelu_erf_fast(v[16 + i * 8 + k]);
                  st_shared_v4(row_s + (((gbase + i) ^ (row & 7)) << 4), pack_h<F16>(o[0], o[1]), pack_h<F16>(o[2], o[3]),
                               pack_h<F16>(o[4], o[5]), pack_h<F16>(o[6], o[7]));
                }
              } else {
                const int cbase = c * 4;               // 32 channels = four 16-byte chunks of the 64-channel box row
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  st_shared_v4(row_s + (((cbase + i) ^ (row & 7)) << 4), pack_h<F16>(v[i * 8 + 0], v[i * 8 + 1]),
                               pack_h<F16>(v[i * 8 + 2], v[i * 8 + 3]), pack_h<F16>(v[i * 8 + 4], v[i * 8 + 5]),
                               pack_h<F16>(v[i * 8 + 6], v[i * 8 + 7]));
              }
            }
          }
          if (geglu && hf == 0) continue;
          fence_proxy_async();          // staging writes -> visible to the TMA engine
          named_bar_sync(2 + grp, 128);
          if (eg == 0) {
            if (geglu) {
              if ((n0 >> 1) < (p.N >> 1)) tma_store_4d(&tmO, stage_out, n0 >> 1, w0, h0, b0);
            } else if (n0 + hf * 64 < p.N) {
              tma_store_4d(&tmO, stage_out, n0 + hf * 64, w0, h0, b0);
            }
            tma_store_commit();
          }
          buf_busy = true;
          if (p.gst != nullptr && eg < 64) {
            // GroupNorm statistics of the staged box: identical granule / lane / row order to the streamed path below
            // (8 lanes per 8-channel granule, rows sub, sub + 8, ...), so the sums are the same bits.  (The box is refilled
            // only after this group's next barrier 4 + grp.)
            constexpr int TPG = 8;
            const int gi = eg / TPG, sub = eg % TPG;     // 8 granules x 8 lanes
            float gs = 0.f, gq = 0.f;
#pragma unroll 4
            for (int rr = sub; rr < BM; rr += TPG) {
              const uint4 u = ld_shared_v4(stage_s + rr * 128 + ((gi ^ (rr & 7)) << 4));
              const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = unpack_h<F16>(uu[i]);
                gs += f.x + f.y;
                gq += f.x * f.x + f.y * f.y;
              }
            }
#pragma unroll
            for (int o = TPG / 2; o > 0; o >>= 1) {
              gs += __shfl_xor_sync(0xffffffffu, gs, o);
              gq += __shfl_xor_sync(0xffffffffu, gq, o);
            }
            if (sub == 0 && n0 + hf * 64 + gi * 8 < p.N) {
              float* dst = p.gst + ((size_t)b * p.gst_slots + p.gst_slot0 + r) * p.gst_ld + (size_t)((n0 + hf * 64) / 8 + gi) * 2;
              dst[0] = gs;
              dst[1] = gq;
            }
          }
        }
        if (eg == 0) LIDM_TRACE(1 + grp, gl, 6);
        b0 += step_b; r += step_r;
        if (r >= p.tiles_per_img) { r -= p.tiles_per_img; ++b0; }
      }
    } else
    for (int tile = tile_first; tile < tile_end; tile += tile_step, ++lt) {
      const int ab = lt & 1;
      int m_tile, n0;
      tile_coords(tile, m_tile, n0);
      const int bt = m_tile / p.tiles_per_img;
      const int b0 = bt * p.bbox;
      const int b = b0 + rb;
      const bool b_ok = b < p.B;          // a multi-sample tile may run past the batch
      const int r = m_tile - bt * p.tiles_per_img;
      const int th = r / p.tiles_w;
      const int h0 = th * p.Hbox, w0 = (r - th * p.tiles_w) * p.Wbox;
      const int h = h0 + hh, w = w0 + ww;
      const int pix = h * p.W + w;
      // stage bias + per-sample row (timestep embedding) for this tile's columns
      for (int i = e; i < BN; i += L::EPI_THREADS) {
        const int n = n0 + i;
        float v = 0.f;
        if (n < p.N) {
          if (p.bias != nullptr) v = __ldg(p.bias + n);
          if (p.rowadd != nullptr) v += __ldg(p.rowadd + (size_t)b0 * p.rowadd_ld + n);
        }
        sbias[ab * SB + i] = v;
      }
      if (use_tma_store && e == 0) tma_store_wait_read<L::OUT_BUFS - 1>();   // this tile's staging buffer is free
      named_bar_sync(1, L::EPI_THREADS);
      const uint32_t taddr_row = tmem_base + ab * L::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
      uint8_t* stage_out = smem + L::OUT_OFF + (L::OUT_BUFS == 2 ? ab : 0) * L::OUT_BUF_BYTES;
      if constexpr (BN >= 64) {
        if (use_tma_store) {
          // Fast path (bf16 output through TMA stores): the residual row is requested before the accumulator is
          // even ready, the whole accumulator row is pulled out of TMEM with back-to-back loads and one wait, and
          // the TMEM buffer is handed back to the MMA warp before any arithmetic or store happens.
          constexpr int HALF = L::EPI_WARPS == 16 ? 32 : (L::COLS > 64 ? 64 : L::COLS);   // columns per register batch
          constexpr int NH = L::COLS / HALF;
          constexpr int NCH = HALF / 32;
          const bool has_res = p.res != nullptr && b_ok;
          const bf16* rp = has_res ? p.res + ((size_t)(b * p.H + h) * p.res_Wp + (w + p.res_hl)) * p.res_ld + n0 + col_lo
                                   : nullptr;
#pragma unroll
          for (int hf = 0; hf < NH; ++hf) {
            uint4 rres[NCH][4];
            if (has_res) {
#pragma unroll
              for (int c = 0; c < NCH; ++c) {
                if (n0 + col_lo + hf * HALF + c * 32 < p.N) {
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    rres[c][i] = __ldg(reinterpret_cast<const uint4*>(rp + hf * HALF + c * 32) + i);
                }
              }
            }
            if (hf == 0) {
              mbar_wait(&tfull_bar[ab], (lt >> 1) & 1);
              tcgen05_fence_after();
            }
            uint32_t raw[NCH][32];
#pragma unroll
            for (int c = 0; c < NCH; ++c) tmem_ld_32x32b_x32(taddr_row + col_lo + hf * HALF + c * 32, raw[c]);
            tmem_ld_wait();
            if (hf == NH - 1) {
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) { if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[ab]), 0)); else mbar_arrive(&tempty_bar[ab]); }
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              const int cc = col_lo + hf * HALF + c * 32;   // column offset inside the tile
              if (n0 + cc < p.N) {
                float v[32];
                const float4* sb4 = reinterpret_cast<const float4*>(sbias + ab * SB + cc);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 t = sb4[j];
                  v[4 * j] = __uint_as_float(raw[c][4 * j]) + t.x;
                  v[4 * j + 1] = __uint_as_float(raw[c][4 * j + 1]) + t.y;
                  v[4 * j + 2] = __uint_as_float(raw[c][4 * j + 2]) + t.z;
                  v[4 * j + 3] = __uint_as_float(raw[c][4 * j + 3]) + t.w;
                }
                if (has_res) {
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    float2 f;
                    const float rs = p.res_scale;
                    f = unpack_h<F16>(rres[c][i].x); v[i * 8 + 0] = fmaf(f.x, rs, v[i * 8 + 0]); v[i * 8 + 1] = fmaf(f.y, rs, v[i * 8 + 1]);
                    f = unpack_h<F16>(rres[c][i].y); v[i * 8 + 2] = fmaf(f.x, rs, v[i * 8 + 2]); v[i * 8 + 3] = fmaf(f.y, rs, v[i * 8 + 3]);
                    f = unpack_h<F16>(rres[c][i].z); v[i * 8 + 4] = fmaf(f.x, rs, v[i * 8 + 4]); v[i * 8 + 5] = fmaf(f.y, rs, v[i * 8 + 5]);
                    f = unpack_h<F16>(rres[c][i].w); v[i * 8 + 6] = fmaf(f.x, rs, v[i * 8 + 6]); v[i * 8 + 7] = fmaf(f.y, rs, v[i * 8 + 7]);
                  }
                }
                if (p.geglu) {
                  // GEGLU (lidm/modules/attention.py:36-44) fused: the weight rows were interleaved at pack time so that
                  // every 32-column chunk holds 16 values followed by their 16 gates; 16 outputs leave per chunk
                  const int oc = cc >> 1;
                  uint8_t* gb = stage_out + (oc >> 6) * (BM * 128) + row * 128;
                  const int gbase = (oc & 63) >> 3;
#pragma unroll
                  for (int i = 0; i < 2; ++i) {
                    float o[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                      const float gt = v[16 + i * 8 + k];
                      o[k] = v[i * 8 + k] * gelu_erf_fast(gt);
                    }
                    uint4 pk;
                    pk.x = pack_h<F16>(o[0], o[1]); pk.y = pack_h<F16>(o[2], o[3]);
                    pk.z = pack_h<F16>(o[4], o[5]); pk.w = pack_h<F16>(o[6], o[7]);
                    *reinterpret_cast<uint4*>(gb + (((gbase + i) ^ (row & 7)) << 4)) = pk;
                  }
                  continue;
                }
                uint8_t* box = stage_out + (cc >> 6) * (BM * 128) + row * 128;
                const int cbase = (cc & 63) >> 3;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  uint4 pk;
                  pk.x = pack_h<F16>(v[i * 8 + 0], v[i * 8 + 1]);
                  pk.y = pack_h<F16>(v[i * 8 + 2], v[i * 8 + 3]);
                  pk.z = pack_h<F16>(v[i * 8 + 4], v[i * 8 + 5]);
                  pk.w = pack_h<F16>(v[i * 8 + 6], v[i * 8 + 7]);
                  *reinterpret_cast<uint4*>(box + (((cbase + i) ^ (row & 7)) << 4)) = pk;
                }
              }
            }
          }
          fence_proxy_async();          // staging writes -> visible to the TMA engine
          named_bar_sync(2, L::EPI_THREADS);
          if (e == 0) {
            if (p.geglu) {
#pragma unroll
              for (int j = 0; j < L::OUT_BOXES / 2; ++j)
                if ((n0 >> 1) + j * 64 < (p.N >> 1))
                  tma_store_4d(&tmO, stage_out + j * (BM * 128), (n0 >> 1) + j * 64, w0, h0, b0);
            } else {
#pragma unroll
              for (int j = 0; j < L::OUT_BOXES; ++j)
                if (n0 + j * 64 < p.N)
                  tma_store_4d(&tmO, stage_out + j * (BM * 128), n0 + j * 64, w0, h0, b0);
            }
            tma_store_commit();
          }
          if (p.gst != nullptr) {
            // GroupNorm statistics of this tile from the staged (bf16-rounded) values: TPG consecutive lanes share one
            // 8-channel granule and walk rows sub, sub + TPG, ...; fixed-order shuffle tree -> deterministic
            // (8 lanes per granule whatever the tile width, so the summation order — and with it the result, bit for
            // bit — does not depend on which BN the launcher picked for this batch size)
            constexpr int GRAN = BN / 8;
            constexpr int TPG = 8;
            const int gi = e / TPG, sub = e % TPG;
            float gs = 0.f, gq = 0.f;
            if (gi < GRAN) {
              const uint8_t* gbox = stage_out + (gi >> 3) * (BM * 128);
#pragma unroll 4
              for (int rr = sub; rr < BM; rr += TPG) {
                const uint4 u = *reinterpret_cast<const uint4*>(gbox + rr * 128 + (((gi & 7) ^ (rr & 7)) << 4));
                const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float2 f = unpack_h<F16>(uu[i]);
                  gs += f.x + f.y;
                  gq += f.x * f.x + f.y * f.y;
                }
              }
            }
#pragma unroll
            for (int o = TPG / 2; o > 0; o >>= 1) {
              gs += __shfl_xor_sync(0xffffffffu, gs, o);
              gq += __shfl_xor_sync(0xffffffffu, gq, o);
            }
            if (sub == 0 && gi < GRAN && n0 + gi * 8 < p.N) {
              float* dst = p.gst + ((size_t)b * p.gst_slots + p.gst_slot0 + r) * p.gst_ld + (size_t)(n0 / 8 + gi) * 2;
              dst[0] = gs;
              dst[1] = gq;
            }
          }
          continue;
        }
      }
      mbar_wait(&tfull_bar[ab], (lt >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = col_lo; c0 < col_lo + L::COLS; c0 += CH) {
        const int n = n0 + c0;
        const bool live = n < p.N && b_ok;   // warp-uniform when bbox == 1
        uint4 rres[4];
        const bool has_res = live && p.res != nullptr && CH == 32;
        if (has_res) {
          const uint4* r4 = reinterpret_cast<const uint4*>(
              p.res + ((size_t)(b * p.H + h) * p.res_Wp + (w + p.res_hl)) * p.res_ld + n);
#pragma unroll
          for (int i = 0; i < 4; ++i) rres[i] = __ldg(r4 + i);
        }
        float v[CH];
        {
          uint32_t raw[CH];
          if constexpr (CH == 32) tmem_ld_32x32b_x32(taddr_row + c0, raw);
          else tmem_ld_32x32b_x16(taddr_row + c0, raw);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(raw[j]);
        }
        if (c0 + CH >= col_lo + L::COLS) {   // last TMEM read of this thread: hand the accumulator back to the MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) { if (PAIR) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[ab]), 0)); else mbar_arrive(&tempty_bar[ab]); }
        }
        if (!live) continue;
        {
          const float4* sb4 = reinterpret_cast<const float4*>(sbias + ab * SB + c0);
#pragma unroll
          for (int j = 0; j < CH / 4; ++j) {
            const float4 t = sb4[j];
            v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
          }
        }
        if (has_res) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float2 f;
            const float rs = p.res_scale;
            f = unpack_h<F16>(rres[i].x); v[i * 8 + 0] = fmaf(f.x, rs, v[i * 8 + 0]); v[i * 8 + 1] = fmaf(f.y, rs, v[i * 8 + 1]);
            f = unpack_h<F16>(rres[i].y); v[i * 8 + 2] = fmaf(f.x, rs, v[i * 8 + 2]); v[i * 8 + 3] = fmaf(f.y, rs, v[i * 8 + 3]);
            f = unpack_h<F16>(rres[i].z); v[i * 8 + 4] = fmaf(f.x, rs, v[i * 8 + 4]); v[i * 8 + 5] = fmaf(f.y, rs, v[i * 8 + 5]);
            f = unpack_h<F16>(rres[i].w); v[i * 8 + 6] = fmaf(f.x, rs, v[i * 8 + 6]); v[i * 8 + 7] = fmaf(f.y, rs, v[i * 8 + 7]);
          }
        } else if (p.res != nullptr) {
          const bf16* rp = p.res + ((size_t)(b * p.H + h) * p.res_Wp + (w + p.res_hl)) * p.res_ld + n;
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (n + j < p.N) v[j] = fmaf(load_h<F16>(rp + j), p.res_scale, v[j]);
        }
        if (p.res_f32 != nullptr) {
          const float* rp = p.res_f32 + ((size_t)b * HW + pix) * p.res_f32_ld + n;
          if (CH == 32 && n + 32 <= p.N) {
            const float4* r4 = reinterpret_cast<const float4*>(rp);
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) {
              const float4 t = __ldg(r4 + i);
              v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (n + j < p.N) v[j] += __ldg(rp + j);
          }
        }

        if (n >= p.split_n) {
          bf16* ot = p.out_t + ((size_t)b * (p.N - p.split_n) + (n - p.split_n)) * HW + pix;
#pragma unroll
          for (int j = 0; j < CH; ++j) store_h<F16>(ot + (size_t)j * HW, v[j]);
        } else if (p.out != nullptr) {
          if constexpr (CH == 32) {
            uint4 pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              pk[i].x = pack_h<F16>(v[i * 8 + 0], v[i * 8 + 1]);
              pk[i].y = pack_h<F16>(v[i * 8 + 2], v[i * 8 + 3]);
              pk[i].z = pack_h<F16>(v[i * 8 + 4], v[i * 8 + 5]);
              pk[i].w = pack_h<F16>(v[i * 8 + 6], v[i * 8 + 7]);
            }
            if (use_tma_store) {
              // staging layout = TMA box [pixel row][64 ch] (128 B rows, 128B swizzle): chunk16 ^= (row & 7)
              uint8_t* box = stage_out + (c0 >> 6) * (BM * 128) + row * 128;
              const int cbase = (c0 & 63) >> 3;   // first 16-byte chunk of this 32-channel group inside the row
#pragma unroll
              for (int i = 0; i < 4; ++i)
                *reinterpret_cast<uint4*>(box + (((cbase + i) ^ (row & 7)) << 4)) = pk[i];
            } else {
              const size_t rowbase = (size_t)(b * p.H + h) * p.out_Wp;
              uint4* o = reinterpret_cast<uint4*>(p.out + (rowbase + w + p.out_hl) * p.out_ld + n);
#pragma unroll
              for (int i = 0; i < 4; ++i) o[i] = pk[i];
              if (w < p.out_hr) {
                uint4* o2 = reinterpret_cast<uint4*>(p.out + (rowbase + p.W + p.out_hl + w) * p.out_ld + n);
#pragma unroll
                for (int i = 0; i < 4; ++i) o2[i] = pk[i];
              }
              if (w >= p.W - p.out_hl) {
                uint4* o2 = reinterpret_cast<uint4*>(p.out + (rowbase + (w - (p.W - p.out_hl))) * p.out_ld + n);
#pragma unroll
                for (int i = 0; i < 4; ++i) o2[i] = pk[i];
              }
            }
          } else {
            bf16* o = p.out + ((size_t)(b * p.H + h) * p.out_Wp + (w + p.out_hl)) * p.out_ld + n;
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (n + j < p.N) store_h<F16>(o + j, v[j]);
          }
        }
        if (p.out_f32_nhwc != nullptr) {
          float* o = p.out_f32_nhwc + ((size_t)b * HW + pix) * p.out_f32_ld + n;
          if (CH == 32 && n + 32 <= p.N) {
            float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
            for (int i = 0; i < CH / 4; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (n + j < p.N) o[j] = v[j];
          }
        }
        if (p.out_f32_nchw != nullptr || p.ddim_x_prev != nullptr) {
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            if (n + j < p.N) {
              const size_t idx = ((size_t)b * p.N + (n + j)) * HW + pix;
              if (p.out_f32_nchw != nullptr) p.out_f32_nchw[idx] = v[j];
              if (p.ddim_x_prev != nullptr) {
                const float nz = p.ddim_noise != nullptr ? p.ddim_noise[idx] : 0.f;
                float xp, x0;
                ddim_update(p.ddim_x[idx], v[j], nz, coef, xp, x0);
                p.ddim_x_prev[idx] = xp;
                if (p.ddim_pred_x0 != nullptr) p.ddim_pred_x0[idx] = x0;
              }
            }
          }
        }
      }
    }
    if (use_tma_store && e == 0) tma_store_wait<0>();   // smem must outlive the in-flight stores
    tcgen05_fence_before();
  }

  if (PAIR) cluster_sync_all();      // the peer's tensor memory and barriers stay valid until both CTAs are done
  else __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, L::TMEM_COLS);
    else tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

template <int BN, int STAGES, bool F16, int RESK = 0, bool PAIR = false>
void launch_persist(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmA2,
                    const GemmKernelParams& p,
                    int num_m_tiles, int num_tiles, int use_tma_store, cudaStream_t stream) {
  using L = PersistLayout<BN, STAGES, RESK, PAIR>;
  static int num_sms = 0;
  if (num_sms == 0) {
    LIDM_CUDA_CHECK(cudaFuncSetAttribute(conv_gemm_persist_kernel<BN, STAGES, F16, RESK, PAIR>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    int dev = 0;
    LIDM_CUDA_CHECK(cudaGetDevice(&dev));
    LIDM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int grid = num_tiles < num_sms ? num_tiles : num_sms;
  if (RESK > 0) {
    // every CTA owns one output-channel tile: a whole number of CTAs per channel tile, no more than there are pixel tiles
    const int num_n_tiles = num_tiles / num_m_tiles;
    int per_n = num_sms / num_n_tiles;
    if (per_n > num_m_tiles) per_n = num_m_tiles;
    grid = per_n * num_n_tiles;
  }
  if (PAIR) {
    grid &= ~1;                        // whole clusters of two
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(L::THREADS); cfg.dynamicSmemBytes = L::TOTAL; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    int na = 1;
    if (pdl_enabled()) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    LIDM_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_gemm_persist_kernel<BN, STAGES, F16, RESK, PAIR>, tmA, tmB, tmO, tmA2, p, num_m_tiles,
                                       num_tiles, use_tma_store));
  } else {
    launch_pdl(conv_gemm_persist_kernel<BN, STAGES, F16, RESK, PAIR>, dim3(grid), dim3(L::THREADS), L::TOTAL, stream, tmA, tmB, tmO, tmA2,
               p, num_m_tiles, num_tiles, use_tma_store);
  }
  LIDM_COUNT_LAUNCH(1);
}

}  // namespace

bool conv_gemm_emits_gstats(const GemmEpilogue& ep, int n_alloc) {
  return ep.out.p != nullptr && ep.out.gst != nullptr && ep.out.hl == 0 && ep.out.hr == 0 && ep.out_t == nullptr &&
         n_alloc % 64 == 0 && ep.res_f32 == nullptr && ep.out_f32_nhwc == nullptr && ep.out_f32_nchw == nullptr;
}

void launch_conv_gemm(const View& a, const ConvTaps& taps, const GemmB& wtb, int N, const GemmEpilogue& ep,
                      cudaStream_t stream) {
  const bf16* wt = wtb.p;
  const int n_alloc = wtb.n_alloc;
  const bool wt_batched = wtb.batch_stride != 0;
  LIDM_REQUIRE(a.p != nullptr && wt != nullptr, "null operand");
  LIDM_REQUIRE(a.C % BK == 0, "Cin must be a multiple of 64 (use the im2col path otherwise)");
  LIDM_REQUIRE(a.ld % 8 == 0 && (reinterpret_cast<uintptr_t>(a.p) & 15) == 0, "activation view must be 16B aligned");
  LIDM_REQUIRE(a.wpitch == 0 && ep.residual.wpitch == 0 && (ep.out.wpitch == 0 || ep.out.hl + ep.out.hr == 0),
               "a row pitch override is supported on halo-free outputs only");
  LIDM_REQUIRE(taps.n >= 1 && taps.n <= 9, "1..9 taps");
  LIDM_REQUIRE(wtb.f16 == a.f16 && (ep.out.p == nullptr || ep.out.f16 == a.f16) &&
                   (ep.residual.p == nullptr || ep.residual.f16 == a.f16) && (ep.a2.p == nullptr || ep.a2.f16 == a.f16) &&
                   (!a.f16 || wtb.nseg == 1),
               "GEMM operands and 2-byte outputs must share one element format (bf16 or fp16)");
  const int nseg = wtb.nseg;
  LIDM_REQUIRE(nseg >= 1 && nseg <= 3 && (nseg != 3 || a.lo_off > 0), "operand-split segments");
  if (conv_halo64_applicable(a, taps, wtb, N, ep)) {      // 64 -> 64 channels: one halo box per tile, resident weights
    launch_conv_halo64(a, taps, wtb, N, ep, stream);
    return;
  }
  const int sx = taps.sx, sy = taps.sy;
  const bool strided = sx != 1 || sy != 1;
  LIDM_REQUIRE(sx >= 1 && sy >= 1 && a.W % sx == 0 && a.H % sy == 0, "stride must divide the input");
  const int W = a.W / sx, H = a.H / sy;      // output pixels (= input pixels unless strided)
  LIDM_REQUIRE(!strided || (taps.cstep == 0 && !taps.zero_w && nseg == 1 && ep.a2.p == nullptr && !wt_batched && H * W >= BM),
               "strided convolution: plain circular conv, one sample per tile");
  LIDM_REQUIRE((W <= BM && BM % W == 0) || (W % BM == 0), "W must divide or be a multiple of 128");
  const int Wbox = W < BM ? W : BM;
  int Hbox = BM / Wbox;
  // a sample smaller than one 128-row tile (the 2x32 level of the layout U-Net): the tile covers `bbox` whole samples
  int bbox = 1;
  if (H < Hbox) {
    LIDM_REQUIRE(Hbox % H == 0 && Hbox / H <= 4, "a sample must hold 32, 64 or a multiple of 128 pixels");
    bbox = Hbox / H;
    Hbox = H;
    LIDM_REQUIRE(ep.rowadd == nullptr || ep.rowadd_ld == 0, "per-sample epilogue rows need one sample per tile");
    LIDM_REQUIRE(!wt_batched && ep.out_t == nullptr, "batched weights need one sample per tile");
  }
  LIDM_REQUIRE(H % Hbox == 0, "H must be a multiple of 128/W");
  for (int t = 0; t < taps.n && taps.cstep == 0 && !taps.zero_w; ++t) {
    LIDM_REQUIRE(-taps.dx[t] <= a.hl && (W - 1) * sx + taps.dx[t] <= a.W - 1 + a.hr, "tap exceeds the materialised halo");
  }
  if (taps.zero_w) LIDM_REQUIRE(a.hl == 0 && a.hr == 0, "zero-padded convolutions read halo-free tensors (TMA fills the border)");
  int BN;
  static const int force_bn = getenv("LIDM_GEMM_BN") ? atoi(getenv("LIDM_GEMM_BN")) : 0;
  // resident weights (PersistLayout RESK): 1x1 GEMMs over 256 or 512 input channels whose output leaves through TMA stores
  // (LIDM_GEMM_RESB = 0: off, 1: both, 256 / 512: that K only - A/B switch).  Measured at B = 64: 256 -> 768 @16x128 71.7 ->
  // 67.1 us resident; 512 -> 1536 @8x64 55.7 -> 67.5 us (twice the tiles at BN = 128, single staging buffer): K = 256 only.
  static const int resb_on = getenv("LIDM_GEMM_RESB") ? atoi(getenv("LIDM_GEMM_RESB")) : 256;
  int resk = 0;
  if (resb_on && !strided && taps.n == 1 && nseg == 1 && ep.a2.p == nullptr && !wt_batched && bbox == 1 && a.C == 256 && ep.residual.p == nullptr &&
      n_alloc % 128 == 0 && n_alloc / 128 <= 148 && ep.out.p != nullptr && ep.out.hl == 0 && ep.out.hr == 0 && ep.out_t == nullptr &&
      ep.res_f32 == nullptr && ep.out_f32_nhwc == nullptr && ep.out_f32_nchw == nullptr && ep.rowadd == nullptr &&
      (resb_on == 1 || a.C == resb_on))   // (a 512-channel variant, B resident in 128 KB with one staging buffer, measured slower)
    resk = a.C;
  if (resk) BN = 128;
  else if (n_alloc % 256 == 0 && force_bn != 128 && ep.out_t == nullptr) {
    // 128x256 tiles halve the B-operand traffic per MAC; take them unless wave quantisation on 148 SMs eats the gain
    const long m_tiles = bbox > 1 ? (a.B + bbox - 1) / bbox : (long)a.B * (H / Hbox) * (W / Wbox);
    auto eff = [&](long tiles) { const long rounds = (tiles + 147) / 148; return (double)tiles / (double)(rounds * 148); };
    const double e256 = eff(m_tiles * (n_alloc / 256)) * 1.30, e128 = eff(m_tiles * (n_alloc / 128));
    BN = (e256 >= e128 || force_bn == 256) ? 256 : 128;
  } else if (n_alloc % 128 == 0) BN = 128;
  else if (n_alloc % 64 == 0) BN = 64;
  else { LIDM_REQUIRE(n_alloc % 16 == 0, "n_alloc must be a multiple of 16"); BN = 16; }
  {
    // small batches: a GEMM that fills less than half of the chip with 128-wide tiles takes 64-wide ones (twice the CTAs,
    // each streaming half of the weight slice: the coarse 3x3 convs at B <= 8 are bound by per-SM weight streaming)
    static const int small_bn = getenv("LIDM_GEMM_SMALL_BN") ? atoi(getenv("LIDM_GEMM_SMALL_BN")) : 1;
    const long m_tiles = bbox > 1 ? (a.B + bbox - 1) / bbox : (long)a.B * (H / Hbox) * (W / Wbox);
    if (small_bn && !resk && BN == 128 && force_bn == 0 && n_alloc % 64 == 0 && m_tiles * (n_alloc / 128) <= 74 && !ep.geglu &&
        taps.n * nseg * (a.C / BK) >= 32)
      BN = 64;
  }
  LIDM_REQUIRE(N <= n_alloc, "N > n_alloc");
  LIDM_REQUIRE(BN == 16 || N % 32 == 0, "N must be a multiple of 32 (or <= 16)");

  GemmKernelParams p{};
  p.tiles_w = W / Wbox;
  p.tiles_per_img = p.tiles_w * (H / Hbox);
  p.Wbox = Wbox; p.Hbox = Hbox; p.hl = a.hl;
  p.sx = sx; p.sy = sy;
  p.bbox = bbox; p.B = a.B;
  p.ntaps = taps.n * nseg;
  for (int t = 0; t < taps.n; ++t) {
    for (int sg = 0; sg < nseg; ++sg) {
      const int vt = t * nseg + sg;
      p.dx[vt] = taps.cstep ? 0 : taps.dx[t];
      p.dy[vt] = taps.cstep ? 0 : taps.dy[t];
      p.a_coff[vt] = t * taps.cstep + ((nseg == 3 && sg == 1) ? a.lo_off : 0);
      p.b_koff[vt] = vt * a.C;
    }
  }
  p.kchunks = a.C / BK;
  p.N = N; p.H = H; p.W = W;
  p.wt_batched = wt_batched ? 1 : 0;
  p.bias = ep.bias; p.rowadd = ep.rowadd; p.rowadd_ld = ep.rowadd_ld;
  if (ep.residual.p != nullptr) {
    LIDM_REQUIRE(ep.residual.H == H && ep.residual.W == W && ep.residual.B == a.B, "residual shape mismatch");
    LIDM_REQUIRE(ep.residual.ld % 8 == 0, "residual ld");
    p.res = ep.residual.p; p.res_ld = ep.residual.ld; p.res_hl = ep.residual.hl; p.res_Wp = ep.residual.Wp();
    p.res_scale = ep.res_scale;
  }
  if (ep.out.p != nullptr) {
    LIDM_REQUIRE(ep.out.H == H && ep.out.W == W && ep.out.B == a.B, "output shape mismatch");
    LIDM_REQUIRE(ep.out.ld % 8 == 0 && (reinterpret_cast<uintptr_t>(ep.out.p) & 15) == 0, "output alignment");
    p.out = ep.out.p; p.out_ld = ep.out.ld; p.out_hl = ep.out.hl; p.out_hr = ep.out.hr; p.out_Wp = ep.out.pitch();
  }
  p.split_n = ep.split_n; p.out_t = ep.out_t;
  if (ep.out_t != nullptr) LIDM_REQUIRE(ep.split_n % BN == 0, "split_n must be tile aligned");
  p.out_f32_nchw = ep.out_f32_nchw; p.out_f32_nhwc = ep.out_f32_nhwc;
  p.out_f32_ld = ep.out_f32_ld ? ep.out_f32_ld : N;
  p.res_f32 = ep.res_f32; p.res_f32_ld = ep.res_f32_ld;
  if (ep.res_f32 != nullptr) LIDM_REQUIRE(ep.out.p == nullptr || ep.out.hl + ep.out.hr > 0 || true, "fp32 residual");
  p.ddim_x = ep.ddim_x; p.ddim_noise = ep.ddim_noise; p.ddim_x_prev = ep.ddim_x_prev;
  p.ddim_pred_x0 = ep.ddim_pred_x0; p.ddim_coef = ep.ddim_coef;

  uint64_t Ktot = (uint64_t)taps.n * nseg * a.C;
  CUtensorMap tmA = make_tma_act(a, BK, Wbox, Hbox, 128, bbox, sx, sy);
  CUtensorMap tmA2 = tmA;
  if (ep.a2.p != nullptr) {
    const View& a2 = ep.a2;
    LIDM_REQUIRE(nseg == 1 && a2.H == H && a2.W == W && a2.B == a.B && a2.C % BK == 0 && a2.ld % 8 == 0 && a2.wpitch == 0 &&
                     (reinterpret_cast<uintptr_t>(a2.p) & 15) == 0 && ep.residual.p == nullptr,
                 "second A operand: same pixels, channels a multiple of 64, bf16 mode, no residual");
    tmA2 = make_tma_act(a2, BK, Wbox, Hbox, 128, bbox);
    p.k2chunks = a2.C / BK; p.hl2 = a2.hl; p.b2_koff = (int)Ktot;
    Ktot += a2.C;
    if (ep.a2_diag) {
      LIDM_REQUIRE(a2.C == N && N % BN == 0 && BN >= BK, "identity-folded residual: a2 must have the output's channels, whole tiles");
      p.a2_diag = 1;
    }
  }
  const uint64_t wld = wtb.ld != 0 ? (uint64_t)wtb.ld : Ktot;
  LIDM_REQUIRE(wld >= Ktot && wld % 8 == 0 && (reinterpret_cast<uintptr_t>(wt) & 15) == 0, "weight operand alignment");
  const int num_m_tiles = bbox > 1 ? (a.B + bbox - 1) / bbox : a.B * p.tiles_per_img;
  const int num_tiles = num_m_tiles * (n_alloc / BN);
  // CTA pairs (cta_group::2) for the wide streamed-weight tiles: two pixel tiles of one channel tile per cluster, each CTA loads
  // half of the B rows (LIDM_GEMM_PAIR=0 switches them off; 128 extends them to the 128-wide tiles, which was measured
  // slower on the decoder's HBM-bound 128-channel levels: (1,4) conv 128 -> 128 @64x512 410 -> 499 us)
  static const int pair_on = getenv("LIDM_GEMM_PAIR") ? atoi(getenv("LIDM_GEMM_PAIR")) : 1;
  const bool pair = pair_on && (BN == 256 || (BN == 128 && pair_on == 128)) && !resk && bbox == 1 && num_m_tiles % 2 == 0 && num_m_tiles >= 2 &&
                    (!wt_batched || p.tiles_per_img % 2 == 0);
  CUtensorMap tmB = make_tma_3d(wt, Ktot, (uint64_t)n_alloc, wt_batched ? (uint64_t)a.B : 1, wld * 2,
                                wt_batched ? (uint64_t)wtb.batch_stride * 2 : wld * 2 * (uint64_t)n_alloc, BK, pair ? BN / 2 : BN,
                                128);
  {
    // Tile order: the output-channel tiles of one pixel tile run back to back (and so concurrently on neighbouring
    // CTAs), so an A tile comes from HBM once and from L2 afterwards.  Measured in the U-Net at B = 64: the wide 3x3
    // convs at the coarse levels gain 5-8 % (1536->1024 @4x32 191 -> 176 us), 1x1 GEMMs 0-6 %, nothing loses.
    // LIDM_GEMM_NFAST=0 restores pixel-tile-fastest order, 1 limits it to 1x1 GEMMs (A/B switch).
    static const int nfast = getenv("LIDM_GEMM_NFAST") ? atoi(getenv("LIDM_GEMM_NFAST")) : 2;
    p.n_fast = (nfast == 2 || (nfast == 1 && p.ntaps == 1)) ? 1 : 0;
  }
  const int use_tma_store = (ep.out.p != nullptr && ep.out.hl == 0 && ep.out.hr == 0 && ep.out_t == nullptr && BN >= 64 &&
                             ep.res_f32 == nullptr && ep.out_f32_nhwc == nullptr && ep.out_f32_nchw == nullptr)
                                ? 1 : 0;
  CUtensorMap tmO = tmA;
  if (use_tma_store) tmO = make_tma_act(ep.out, 64, Wbox, Hbox, 128, bbox);
  if (ep.geglu) {
    LIDM_REQUIRE(use_tma_store && BN >= 128 && N % 128 == 0 && ep.out.C == N / 2 && ep.residual.p == nullptr && ep.out.gst == nullptr,
                 "GEGLU epilogue needs a bf16 TMA-store output of N/2 channels and 128-column tiles");
    p.geglu = 1;
  }
  if (use_tma_store && ep.out.gst != nullptr) {
    LIDM_REQUIRE(bbox == 1 && Wbox * Hbox == 128 && N % 8 == 0 && ep.out.gst_slots >= ep.out.gst_slot0 + p.tiles_per_img,
                 "GroupNorm statistics need whole 128-pixel tiles inside one sample");
    p.gst = ep.out.gst; p.gst_ld = ep.out.gst_ld; p.gst_slots = ep.out.gst_slots; p.gst_slot0 = ep.out.gst_slot0;
  }
  static const char* trace_path = getenv("LIDM_GEMM_TRACE");
  static long long* trace_dev = nullptr;
  static int trace_left = getenv("LIDM_GEMM_TRACE_N") ? atoi(getenv("LIDM_GEMM_TRACE_N")) : 4;   // dump the first few qualifying launches
  const bool tracing = LIDM_GEMM_TRACE_ON && trace_path != nullptr && resk != 0 && trace_left > 0;
  if (tracing) {
    if (trace_dev == nullptr) LIDM_CUDA_CHECK(cudaMalloc(&trace_dev, 8 * 64 * 8 * sizeof(long long)));
    LIDM_CUDA_CHECK(cudaMemsetAsync(trace_dev, 0, 8 * 64 * 8 * sizeof(long long), stream));
    p.trace = trace_dev;
  }
#define LIDM_LAUNCH_GEMM(F16)                                                                                         \
  do {                                                                                                                  \
    if (resk == 256) launch_persist<128, 8, F16, 256>(tmA, tmB, tmO, tmA2, p, num_m_tiles, num_tiles, use_tma_store, stream); \
    else if (BN == 256 && pair) launch_persist<256, 4, F16, 0, true>(tmA, tmB, tmO, tmA2, p, num_m_tiles, num_tiles, use_tma_store, stream); \
    else if (BN == 256) launch_persist<256, 3, F16>(tmA, tmB, tmO, tmA2, p, num_m_tiles, num_tiles, use_tma_store, stream);   \
    else if (BN == 128 && pair) launch_persist<128, 6, F16, 0, true>(tmA, tmB, tmO, tmA2, p, num_m_tiles, num_tiles, use_tma_store, stream); \
    else if (BN == 128) launch_persist<128, 4, F16>(tmA, tmB, tmO, tmA2, p, num_m_tiles, num_tiles, use_tma_store, stream); \
    else if (BN == 64) launch_persist<64, 6, F16>(tmA, tmB, tmO, tmA2, p, num_m_tiles, num_tiles, use_tma_store, stream); \
    else launch_persist<16, 6, F16>(tmA, tmB, tmO, tmA2, p, num_m_tiles, num_tiles, use_tma_store, stream);             \
  } while (0)
  if (a.f16) LIDM_LAUNCH_GEMM(true);
  else LIDM_LAUNCH_GEMM(false);
#undef LIDM_LAUNCH_GEMM
  if (tracing) {
    std::vector<long long> hbuf(8 * 64 * 8);
    LIDM_CUDA_CHECK(cudaStreamSynchronize(stream));
    LIDM_CUDA_CHECK(cudaMemcpy(hbuf.data(), trace_dev, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    FILE* f = fopen(trace_path, "a");
    if (f != nullptr) {
      fprintf(f, "# launch K=%d N=%d m_tiles=%d res=%d gst=%d  CTA 0; row 0 = producer / MMA: A_issue A_issued mma_tile_start acc_free tile_landed - committed; rows 1, 2 = epilogue groups: start tfull - acc_released - - tile_done\n",
              a.C, N, num_m_tiles, ep.residual.p != nullptr, p.gst != nullptr);
      for (int c = 0; c < 4; ++c)
        for (int t = 0; t < 64; ++t) {
          const long long* r = &hbuf[((size_t)c * 64 + t) * 8];
          if (r[0] == 0) break;
          const long long t0 = hbuf[0];
          fprintf(f, "%d %d", c, t);
          for (int k = 0; k < 8; ++k) fprintf(f, " %lld", r[k] - t0);
          fprintf(f, "\n");
        }
      fclose(f);
    }
    --trace_left;
  }
}

}  // namespace lidm
