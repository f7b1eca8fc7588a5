// Halo-tile convolution for 64-channel tensors on tcgen05 / TMEM / TMA (sm_100a).
//
// The full-resolution levels of the R2DM U-Net (efficient_unet.py: 64 -> 64 ring convs at 64x1024) and of the first-stage
// autoencoder (model_lidm.py: (1,4) / 3x3 convs over 64 channels at 64x1024) are GEMMs with N = 64 and one 64-channel K chunk per
// tap.  On the streaming kernel (gemm_conv.cu) such a tile re-loads its activation box for every tap (9 x 16 KB for 16 KB of
// output) and pays the issuing thread's per-chunk barrier / descriptor work for every four N = 64 MMAs - the issue loop, not
// the tensor pipe, sets the pace (~3600 clocks per tile against 1728 of MMA).  Here
//   * the weights (taps x 64 rows x 128 B) stay in shared memory for the life of the CTA,
//   * ONE TMA box per tile brings the pixel tile WITH its halo - rows dy_min..dy_max, 136 pixels (128 + taps' dx range,
//     padded to whole 8-row swizzle atoms) - and every tap's A operand is a descriptor into that box: start address shifted by
//     (dy - dy_min) * 136 * 128 + (dx - dx_min) * 128 bytes, the descriptor's base-offset field carrying the phase of the start
//     row inside the 1024-byte SWIZZLE_128B atom,
//   * so a tile is one barrier wait and taps x 4 back-to-back tcgen05.mma,
//   * two epilogue groups of four warps take alternate tiles (accumulator g, staging box g), as in the resident-weight GEMM.
// Epilogue: bias, optional scaled residual, 2-byte output through a TMA store, GroupNorm granule statistics (same granule /
// lane / row order as gemm_conv.cu, so the statistics keep their bits whichever kernel produced the tensor).
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace lidm {

namespace {

constexpr int HT_BM = 128;                  // pixels per tile (one row segment)
constexpr int HT_N = 64;                    // output channels
constexpr int HT_PX = 136;                  // pixels of the halo box per row (17 swizzle atoms of 8 rows)
constexpr int HT_ROW_BYTES = HT_PX * 128;   // one halo row of 64 channels
constexpr int HT_MAX_TAPS = 9;
constexpr int HT_MAX_ROWS = 3;

struct HaloParams {
  int ntaps, nrows;            // taps, halo rows (dy range)
  int dxmin, dymin;
  int8_t dx[HT_MAX_TAPS], dy[HT_MAX_TAPS];
  int tiles_w, tiles_per_img, num_tiles, hl;
  int N, H, W;
  const float* bias;
  const bf16* res;
  float res_scale;
  int res_ld, res_hl, res_Wp;
  float* gst;
  int gst_ld, gst_slots, gst_slot0;
};

struct HaloLayout {
  static constexpr int B_OFF = 0;                                       // [taps][64 rows][128 B]
  static constexpr int B_TAP = HT_N * 128;                              // 8 KB
  static constexpr int A_OFF = HT_MAX_TAPS * B_TAP;                     // two halo tiles
  static constexpr int A_TILE = HT_MAX_ROWS * HT_ROW_BYTES;             // 52224
  static constexpr int OUT_OFF = A_OFF + 2 * A_TILE;
  static constexpr int OUT_BUF = HT_BM * 128;                           // one 64-channel box per group
  static constexpr int BIAS_OFF = OUT_OFF + 2 * OUT_BUF;
  static constexpr int BAR_OFF = BIAS_OFF + HT_N * 4;
  static constexpr int TOTAL = BAR_OFF + 128 + 1024;
  static_assert(A_OFF % 1024 == 0 && A_TILE % 1024 == 0 && OUT_OFF % 1024 == 0, "swizzle atoms must stay 1024-byte aligned");
  static_assert(TOTAL <= 227 * 1024, "shared memory budget");
};

// A tap's A operand starts (dx - dxmin) rows into a 1024-byte swizzle atom of the halo tile.  The descriptor is the plain K-major
// SWIZZLE_128B one with that start address and base-offset bits 49-51 left ZERO: the tensor core applies the 128-byte swizzle to the
// absolute shared-memory address (bits 4-6 ^= bits 7-9), exactly as TMA did when it wrote the tile, so a start address in the
// middle of an atom already carries its phase.  (Measured: setting the base offset to (addr >> 7) & 7 shifts the phase a second
// time and the convolution comes out wrong - rel 0.6-0.8 in tests/test_gpu_ops.py::test_halo_tile_conv_matches_streamed_kernel.)

template <bool F16>
__global__ void __launch_bounds__(320, 1)
conv_halo64_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmO, const HaloParams p) {
  using L = HaloLayout;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* afull = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);    // [2] halo tile landed
  uint64_t* aempty = afull + 2;                                        // [2] its MMAs have completed
  uint64_t* tfull = aempty + 2;                                        // [2] accumulator ready
  uint64_t* tempty = tfull + 2;                                        // [2] accumulator drained
  uint64_t* bres = tempty + 2;                                         // weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bres + 1);
  float* sbias = reinterpret_cast<float*>(smem + L::BIAS_OFF);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); prefetch_tensormap(&tmO);
    for (int i = 0; i < 2; ++i) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    mbar_init(bres, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  const int tile_first = blockIdx.x, tile_step = gridDim.x;
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    const bool leader = elect_one() != 0;
    if (tile_first < p.num_tiles) {
      if (leader) {
        mbar_arrive_expect_tx(bres, (uint32_t)(p.ntaps * L::B_TAP));
        for (int t = 0; t < p.ntaps; ++t) tma_load_3d(smem + L::B_OFF + t * L::B_TAP, &tmB, bres, t * 64, 0, 0);
      }
      __syncwarp();
    }
    int lt = 0;
    for (int tile = tile_first; tile < p.num_tiles; tile += tile_step, ++lt) {
      const int tb = lt & 1;
      const int b = tile / p.tiles_per_img, r = tile - b * p.tiles_per_img;
      const int h = r / p.tiles_w, w0 = (r - h * p.tiles_w) * HT_BM;
      mbar_wait(&aempty[tb], ((lt >> 1) & 1) ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(&afull[tb], (uint32_t)(p.nrows * HT_ROW_BYTES));
        tma_load_4d(smem + L::A_OFF + tb * L::A_TILE, &tmA, &afull[tb], 0, w0 + p.hl + p.dxmin, h + p.dymin, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const bool leader = elect_one() != 0;
    constexpr uint32_t idesc = make_idesc_h<F16>(HT_BM, HT_N);
    if (tile_first < p.num_tiles) mbar_wait(bres, 0);
    int lt = 0;
    for (int tile = tile_first; tile < p.num_tiles; tile += tile_step, ++lt) {
      const int ab = lt & 1;
      mbar_wait(&tempty[ab], ((lt >> 1) & 1) ^ 1);
      mbar_wait(&afull[ab], (lt >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t d = tmem_base + ab * HT_N;
      const uint32_t a_base = smem_u32(smem + L::A_OFF + ab * L::A_TILE);
      const uint32_t b_base = smem_u32(smem + L::B_OFF);
      if (leader) {
        for (int t = 0; t < p.ntaps; ++t) {
          const uint32_t a_addr = a_base + (p.dy[t] - p.dymin) * HT_ROW_BYTES + (p.dx[t] - p.dxmin) * 128;
          const uint64_t adesc = make_kmajor_desc<128>(a_addr);
          const uint64_t bdesc = make_kmajor_desc<128>(b_base + t * L::B_TAP);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(d, adesc + 2 * k, bdesc + 2 * k, idesc, (t | k) != 0);
        }
        umma_commit(&aempty[ab]);
        umma_commit(&tfull[ab]);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue: two groups of four warps, alternate tiles
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int e = threadIdx.x - 64, eg = e & 127;
    const int row = q * 32 + lane;
    for (int i = e; i < HT_N; i += 256) sbias[i] = (i < p.N && p.bias != nullptr) ? __ldg(p.bias + i) : 0.f;
    named_bar_sync(1, 256);
    const uint32_t sbias_s = smem_u32(sbias);
    uint8_t* stage_out = smem + L::OUT_OFF + grp * L::OUT_BUF;
    const uint32_t stage_s = smem_u32(stage_out), row_s = stage_s + row * 128;
    const uint32_t taddr_row = tmem_base + grp * HT_N + (static_cast<uint32_t>(q * 32) << 16);
    int gl = 0;
    for (int tile = tile_first + grp * tile_step; tile < p.num_tiles; tile += 2 * tile_step, ++gl) {
      const int b = tile / p.tiles_per_img, r = tile - b * p.tiles_per_img;
      const int h = r / p.tiles_w, w0 = (r - h * p.tiles_w) * HT_BM;
      uint4 rres[8];
      const bool has_res = p.res != nullptr;
      if (has_res) {
        const uint4* r4 = reinterpret_cast<const uint4*>(p.res + ((size_t)(b * p.H + h) * p.res_Wp + (w0 + row + p.res_hl)) * p.res_ld);
#pragma unroll
        for (int i = 0; i < 8; ++i) rres[i] = __ldg(r4 + i);
      }
      mbar_wait(&tfull[grp], gl & 1);
      tcgen05_fence_after();
      uint32_t raw[2][32];
#pragma unroll
      for (int c = 0; c < 2; ++c) tmem_ld_32x32b_x32(taddr_row + c * 32, raw[c]);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[grp]);
      if (gl > 0) {
        if (eg == 0) tma_store_wait_read<0>();        // this group's previous store has left the staging box
        named_bar_sync(4 + grp, 128);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = ld_shared_f4(sbias_s + (c * 32 + 4 * j) * 4);
          v[4 * j] = __uint_as_float(raw[c][4 * j]) + t.x;
          v[4 * j + 1] = __uint_as_float(raw[c][4 * j + 1]) + t.y;
          v[4 * j + 2] = __uint_as_float(raw[c][4 * j + 2]) + t.z;
          v[4 * j + 3] = __uint_as_float(raw[c][4 * j + 3]) + t.w;
        }
        if (has_res) {
          const float rs = p.res_scale;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 u = rres[c * 4 + i];
            float2 f;
            f = unpack_h<F16>(u.x); v[i * 8 + 0] = fmaf(f.x, rs, v[i * 8 + 0]); v[i * 8 + 1] = fmaf(f.y, rs, v[i * 8 + 1]);
            f = unpack_h<F16>(u.y); v[i * 8 + 2] = fmaf(f.x, rs, v[i * 8 + 2]); v[i * 8 + 3] = fmaf(f.y, rs, v[i * 8 + 3]);
            f = unpack_h<F16>(u.z); v[i * 8 + 4] = fmaf(f.x, rs, v[i * 8 + 4]); v[i * 8 + 5] = fmaf(f.y, rs, v[i * 8 + 5]);
            f = unpack_h<F16>(u.w); v[i * 8 + 6] = fmaf(f.x, rs, v[i * 8 + 6]); v[i * 8 + 7] = fmaf(f.y, rs, v[i * 8 + 7]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          st_shared_v4(row_s + (((c * 4 + i) ^ (row & 7)) << 4), pack_h<F16>(v[i * 8 + 0], v[i * 8 + 1]),
                       pack_h<F16>(v[i * 8 + 2], v[i * 8 + 3]), pack_h<F16>(v[i * 8 + 4], v[i * 8 + 5]),
                       pack_h<F16>(v[i * 8 + 6], v[i * 8 + 7]));
      }
      fence_proxy_async();
      named_bar_sync(2 + grp, 128);
      if (eg == 0) {
        tma_store_4d(&tmO, stage_out, 0, w0, h, b);
        tma_store_commit();
      }
      if (p.gst != nullptr && eg < 64) {
        // GroupNorm statistics of the staged box: 8 lanes per 8-channel granule, rows sub, sub + 8, ... (order of gemm_conv.cu)
        const int gi = eg >> 3, sub = eg & 7;
        float gs = 0.f, gq = 0.f;
#pragma unroll 4
        for (int rr = sub; rr < HT_BM; rr += 8) {
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
        for (int o = 4; o > 0; o >>= 1) {
          gs += __shfl_xor_sync(0xffffffffu, gs, o);
          gq += __shfl_xor_sync(0xffffffffu, gq, o);
        }
        if (sub == 0 && gi * 8 < p.N) {
          float* dst = p.gst + ((size_t)b * p.gst_slots + p.gst_slot0 + r) * p.gst_ld + (size_t)gi * 2;
          dst[0] = gs;
          dst[1] = gq;
        }
      }
    }
    if (eg == 0) tma_store_wait<0>();
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem_base, 128); }
}

}  // namespace

static int g_halo_override = -1;
void conv_halo64_override(int on) { g_halo_override = on; }

bool conv_halo64_applicable(const View& a, const ConvTaps& taps, const GemmB& wtb, int N, const GemmEpilogue& ep) {
  static const int env_on = getenv("LIDM_GEMM_HALO") ? atoi(getenv("LIDM_GEMM_HALO")) : 1;
  const int on = g_halo_override >= 0 ? g_halo_override : env_on;
  if (!on || a.C != 64 || N != 64 || wtb.n_alloc != 64 || wtb.nseg != 1 || wtb.batch_stride != 0 || taps.cstep != 0 || taps.sx != 1 || taps.sy != 1 || taps.n < 2 ||
      taps.n > HT_MAX_TAPS || a.W % HT_BM != 0 || a.wpitch != 0 || a.lo_off != 0 || a.ld != 64)
    return false;
  if (ep.a2.p != nullptr || ep.rowadd != nullptr || ep.out_t != nullptr || ep.geglu || ep.res_f32 != nullptr ||
      ep.out_f32_nchw != nullptr || ep.out_f32_nhwc != nullptr || ep.ddim_x_prev != nullptr || ep.out.p == nullptr ||
      ep.out.hl != 0 || ep.out.hr != 0 || ep.out.wpitch != 0 || ep.out.f16 != a.f16 || wtb.f16 != a.f16)
    return false;
  if (ep.residual.p != nullptr && (ep.residual.ld != 64 || ep.residual.f16 != a.f16 || ep.residual.wpitch != 0)) return false;
  int dxmin = 127, dxmax = -128, dymin = 127, dymax = -128;
  for (int t = 0; t < taps.n; ++t) {
    dxmin = taps.dx[t] < dxmin ? taps.dx[t] : dxmin; dxmax = taps.dx[t] > dxmax ? taps.dx[t] : dxmax;
    dymin = taps.dy[t] < dymin ? taps.dy[t] : dymin; dymax = taps.dy[t] > dymax ? taps.dy[t] : dymax;
  }
  if (dxmax - dxmin > HT_PX - HT_BM || dymax - dymin + 1 > HT_MAX_ROWS) return false;
  if (!taps.zero_w && (-dxmin > a.hl || dxmax > a.hr)) return false;
  if ((reinterpret_cast<uintptr_t>(a.p) & 15) != 0 || (reinterpret_cast<uintptr_t>(ep.out.p) & 15) != 0) return false;
  if (ep.out.gst != nullptr && (ep.out.gst_slots < ep.out.gst_slot0 + (a.H * a.W) / HT_BM)) return false;
  return true;
}

void launch_conv_halo64(const View& a, const ConvTaps& taps, const GemmB& wtb, int N, const GemmEpilogue& ep, cudaStream_t stream) {
  LIDM_REQUIRE(conv_halo64_applicable(a, taps, wtb, N, ep), "halo-tile convolution: unsupported shape");
  HaloParams p{};
  p.ntaps = taps.n;
  int dxmin = 127, dymin = 127, dymax = -128;
  for (int t = 0; t < taps.n; ++t) {
    p.dx[t] = taps.dx[t]; p.dy[t] = taps.dy[t];
    dxmin = taps.dx[t] < dxmin ? taps.dx[t] : dxmin;
    dymin = taps.dy[t] < dymin ? taps.dy[t] : dymin; dymax = taps.dy[t] > dymax ? taps.dy[t] : dymax;
  }
  p.dxmin = dxmin; p.dymin = dymin; p.nrows = dymax - dymin + 1;
  p.tiles_w = a.W / HT_BM; p.tiles_per_img = p.tiles_w * a.H; p.num_tiles = a.B * p.tiles_per_img; p.hl = a.hl;
  p.N = N; p.H = a.H; p.W = a.W;
  p.bias = ep.bias;
  if (ep.residual.p != nullptr) {
    LIDM_REQUIRE(ep.residual.H == a.H && ep.residual.W == a.W && ep.residual.B == a.B, "residual shape mismatch");
    p.res = ep.residual.p; p.res_ld = ep.residual.ld; p.res_hl = ep.residual.hl; p.res_Wp = ep.residual.Wp(); p.res_scale = ep.res_scale;
  }
  if (ep.out.gst != nullptr) { p.gst = ep.out.gst; p.gst_ld = ep.out.gst_ld; p.gst_slots = ep.out.gst_slots; p.gst_slot0 = ep.out.gst_slot0; }
  CUtensorMap tmA = make_tma_act(a, 64, HT_PX, p.nrows, 128);
  const uint64_t Ktot = (uint64_t)taps.n * 64;
  const uint64_t wld = wtb.ld != 0 ? (uint64_t)wtb.ld : Ktot;
  CUtensorMap tmB = make_tma_3d(wtb.p, Ktot, 64, 1, wld * 2, wld * 2 * 64, 64, 64, 128);
  CUtensorMap tmO = make_tma_act(ep.out, 64, HT_BM, 1, 128);
  static int num_sms = 0;
  if (num_sms == 0) {
    LIDM_CUDA_CHECK(cudaFuncSetAttribute(conv_halo64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloLayout::TOTAL));
    LIDM_CUDA_CHECK(cudaFuncSetAttribute(conv_halo64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HaloLayout::TOTAL));
    int dev = 0;
    LIDM_CUDA_CHECK(cudaGetDevice(&dev));
    LIDM_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  if (a.f16) launch_pdl(conv_halo64_kernel<true>, dim3(grid), dim3(320), HaloLayout::TOTAL, stream, tmA, tmB, tmO, p);
  else launch_pdl(conv_halo64_kernel<false>, dim3(grid), dim3(320), HaloLayout::TOTAL, stream, tmA, tmB, tmO, p);
  LIDM_COUNT_LAUNCH(1);
}

}  // namespace lidm
