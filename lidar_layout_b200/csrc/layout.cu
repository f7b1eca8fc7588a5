// Layout-conditioned denoiser pieces (sm_100a): the object-aware cross-attention of LayoutDiffusionUNetModel and the
// small kernels around it.  Reference: lidm/modules/unets/object_cross_unet.py
//   ObjectAwareCrossAttention.forward  :447-565   image tokens attend to image tokens AND the 13 layout tokens;
//                                                 queries / keys = [content (64) | positional (64)] per head, values content
//   ResBlock up / down sampling        :112-171, 253-283   nearest x2 / 2x2 average pooling of h and x
//
//   oaca_attention_kernel   one CTA per (128-query tile, head, sample): the whole key set (T <= 256 image keys + 16 layout
//                           slots) fits one pass, so there is no online softmax: S = Qc Kc^T + Qp Kp^T accumulates in
//                           TMEM (tcgen05.mma, K = 64 + 64, TMA-fed SWIZZLE_128B operands), every softmax thread owns one
//                           row (two passes over its TMEM lane: max, then exp2 / sum / bf16 pack), P goes back to TMEM
//                           (tcgen05.st) and is the A operand of the TS-form P*V product (V = MN-major B operand), O is
//                           normalised and stored as 128-byte rows.  The 128^-1/4 scale is folded into q, k and the
//                           positional tensors upstream.
//   oaca_pos_kernel / oaca_layout_kv_kernel   once per conditioning: the positional projections + GroupNorm32 and the
//                           layout tokens' keys / values (fp32 CUDA cores; 13 tokens, batch-broadcast aware)
//   avgpool2_kernel         F.avg_pool2d(2, 2) on channels-last 2-byte tensors
#include "common.h"
#include "ptx.cuh"

namespace lidm {

namespace {

constexpr int DH = 64;            // head channels (content); queries / keys carry DH positional channels on top
constexpr int LAY = 16;           // layout key slots (13 objects in the shipped configuration, zero-padded)
constexpr int QROWS = 128;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

template <int T>
struct OacaCfg {
  static_assert(T == 64 || T == 128 || T == 192 || T == 256, "image key count");
  static constexpr int OFF_QC = 0;                                   // [128][64] bf16, 128-byte rows, SWIZZLE_128B
  static constexpr int OFF_QP = OFF_QC + QROWS * 128;
  static constexpr int OFF_KC = OFF_QP + QROWS * 128;                // [T][64]
  static constexpr int OFF_KP = OFF_KC + T * 128;
  static constexpr int OFF_V = OFF_KP + T * 128;                     // [T][64] (MN-major B operand of P*V)
  static constexpr int OFF_LC = OFF_V + T * 128;                     // layout keys, content half [16][64]
  static constexpr int OFF_LP = OFF_LC + LAY * 128;                  // layout keys, positional half
  static constexpr int OFF_LV = OFF_LP + LAY * 128;                  // layout values
  static constexpr int OFF_BAR = OFF_LV + LAY * 128;
  static constexpr int SMEM_TOTAL = OFF_BAR + 128 + 1024;
  static constexpr uint32_t S_COLS = T + LAY;
  static constexpr uint32_t P_COL = (S_COLS + 31) / 32 * 32;          // bf16 pairs: (T + 16) / 2 columns
  static constexpr uint32_t O_COL = (P_COL + S_COLS / 2 + 31) / 32 * 32;
  static constexpr uint32_t TMEM_COLS = (O_COL + DH) <= 256 ? 256 : 512;
  static_assert(O_COL + DH <= 512, "TMEM budget");
  static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");
};

// qkv: (B, T, 3C) = [q | k | v] (q, k pre-scaled); pos: (Bp, T, C) positional half of queries and image keys (pre-scaled);
// klay: (B, 16, 2C) = [layout key content | layout key positional]; vlay: (B, 16, C); out (B, T, C).
template <int T, bool F16>
__global__ void __launch_bounds__(160, 1)
oaca_attention_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmPos,
                      const __grid_constant__ CUtensorMap tmKL, const __grid_constant__ CUtensorMap tmVL,
                      bf16* __restrict__ out, int out_ld, int C, int pos_batched, int n_layout) {
  using L = OacaCfg<T>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_qk = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* full_v = full_qk + 1;
  uint64_t* s_ready = full_qk + 2;
  uint64_t* p_ready = full_qk + 3;
  uint64_t* o_ready = full_qk + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_qk + 5);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // warp index: warp-uniform for the compiler
  const int q0 = blockIdx.x * QROWS, head = blockIdx.y, b = blockIdx.z;
  constexpr int QBOX = T < QROWS ? T : QROWS;          // query rows actually loaded (T = 64: half a tile)

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmQKV); prefetch_tensormap(&tmPos); prefetch_tensormap(&tmKL); prefetch_tensormap(&tmVL);
    mbar_init(full_qk, 1); mbar_init(full_v, 1); mbar_init(s_ready, 1); mbar_init(p_ready, 4); mbar_init(o_ready, 1);
    fence_barrier_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, L::TMEM_COLS); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // the whole warp runs this role (warp-uniform operands stay in uniform registers); one elected lane issues
    const bool leader = elect_one() != 0;
    {
      const int pb = pos_batched ? b : 0;
      // ---- loads: 64-row boxes of 64 channels (8 KiB each)
      if (leader) {
      mbar_arrive_expect_tx(full_qk, 2 * QBOX * 128 + 2 * T * 128 + 2 * LAY * 128);
      for (int r = 0; r < QBOX; r += 64) {
        tma_load_3d(smem + L::OFF_QC + r * 128, &tmQKV, full_qk, head * DH, q0 + r, b);
        tma_load_3d(smem + L::OFF_QP + r * 128, &tmPos, full_qk, head * DH, q0 + r, pb);
      }
      for (int r = 0; r < T; r += 64) {
        tma_load_3d(smem + L::OFF_KC + r * 128, &tmQKV, full_qk, C + head * DH, r, b);
        tma_load_3d(smem + L::OFF_KP + r * 128, &tmPos, full_qk, head * DH, r, pb);
      }
      tma_load_3d(smem + L::OFF_LC, &tmKL, full_qk, head * DH, 0, b);
      tma_load_3d(smem + L::OFF_LP, &tmKL, full_qk, C + head * DH, 0, b);
      mbar_arrive_expect_tx(full_v, T * 128 + LAY * 128);
      for (int r = 0; r < T; r += 64) tma_load_3d(smem + L::OFF_V + r * 128, &tmQKV, full_v, 2 * C + head * DH, r, b);
      tma_load_3d(smem + L::OFF_LV, &tmVL, full_v, head * DH, 0, b);
      }
      __syncwarp();

      // ---- S = Qc Kc^T + Qp Kp^T  (image keys: N = T, layout keys: N = 16)
      constexpr uint32_t idesc_img = make_idesc_h<F16>(QROWS, T);
      constexpr uint32_t idesc_lay = make_idesc_h<F16>(QROWS, LAY);
      mbar_wait(full_qk, 0);
      tcgen05_fence_after();
      const uint64_t qc = make_kmajor_desc<128>(smem_u32(smem + L::OFF_QC)), qp = make_kmajor_desc<128>(smem_u32(smem + L::OFF_QP));
      const uint64_t kc = make_kmajor_desc<128>(smem_u32(smem + L::OFF_KC)), kp = make_kmajor_desc<128>(smem_u32(smem + L::OFF_KP));
      const uint64_t lc = make_kmajor_desc<128>(smem_u32(smem + L::OFF_LC)), lp = make_kmajor_desc<128>(smem_u32(smem + L::OFF_LP));
      if (leader) {
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(tmem_base, qc + 2 * k, kc + 2 * k, idesc_img, k != 0);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(tmem_base, qp + 2 * k, kp + 2 * k, idesc_img, 1);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(tmem_base + T, qc + 2 * k, lc + 2 * k, idesc_lay, k != 0);
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) umma_bf16_ss(tmem_base + T, qp + 2 * k, lp + 2 * k, idesc_lay, 1);
      umma_commit(s_ready);
      }
      __syncwarp();

      // ---- O = P V (P from tensor memory, V an MN-major shared-memory operand: 16 keys = two 8-row groups = 2 KiB)
      constexpr uint32_t idesc_o = make_idesc_h<F16>(QROWS, DH) | (1u << 16);
      mbar_wait(full_v, 0);
      mbar_wait(p_ready, 0);
      tcgen05_fence_after();
      const uint64_t vd = make_kmajor_desc<128>(smem_u32(smem + L::OFF_V)), lv = make_kmajor_desc<128>(smem_u32(smem + L::OFF_LV));
      const uint32_t tP = tmem_base + L::P_COL, dO = tmem_base + L::O_COL;
      if (leader) {
#pragma unroll
      for (int kk = 0; kk < T / 16; ++kk) umma_bf16_ts(dO, tP + kk * 8, vd + (uint64_t)((kk * 2048) >> 4), idesc_o, kk != 0);
      umma_bf16_ts(dO, tP + (T / 16) * 8, lv, idesc_o, 1);
      umma_commit(o_ready);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ softmax: one thread per query row
    const int row = warp * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    mbar_wait(s_ready, 0);
    tcgen05_fence_after();
    float m = -INFINITY;
    {
#pragma unroll 1
      for (int c = 0; c < T / 32; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32b_x32(lane_base + c * 32, sv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(sv[i]));
      }
      uint32_t lv16[16];
      tmem_ld_32x32b_x16(lane_base + T, lv16);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < LAY; ++i)
        if (i < n_layout) m = fmaxf(m, __uint_as_float(lv16[i]));
    }
    const float mb = m * LOG2E;
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < T / 64; ++c) {
      uint32_t sa[32], sb[32], pk[32];
      tmem_ld_32x32b_x32(lane_base + c * 64, sa);
      tmem_ld_32x32b_x32(lane_base + c * 64 + 32, sb);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = ex2f(fmaf(__uint_as_float(sa[2 * i]), LOG2E, -mb)), p1 = ex2f(fmaf(__uint_as_float(sa[2 * i + 1]), LOG2E, -mb));
        const float p2 = ex2f(fmaf(__uint_as_float(sb[2 * i]), LOG2E, -mb)), p3 = ex2f(fmaf(__uint_as_float(sb[2 * i + 1]), LOG2E, -mb));
        sum += (p0 + p1) + (p2 + p3);
        pk[i] = pack_h<F16>(p0, p1);
        pk[16 + i] = pack_h<F16>(p2, p3);
      }
      tmem_st_32x32b_x32(lane_base + L::P_COL + c * 32, pk);
    }
    {
      uint32_t lv16[16], pk[8];
      tmem_ld_32x32b_x16(lane_base + T, lv16);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float p0 = (2 * i < n_layout) ? ex2f(fmaf(__uint_as_float(lv16[2 * i]), LOG2E, -mb)) : 0.f;
        const float p1 = (2 * i + 1 < n_layout) ? ex2f(fmaf(__uint_as_float(lv16[2 * i + 1]), LOG2E, -mb)) : 0.f;
        sum += p0 + p1;
        pk[i] = pack_h<F16>(p0, p1);
      }
      tmem_st_32x32b_x8(lane_base + L::P_COL + T / 2, pk);
    }
    tmem_st_wait();
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_ready);
    // ---- epilogue
    mbar_wait(o_ready, 0);
    tcgen05_fence_after();
    uint32_t o0[32], o1[32];
    tmem_ld_32x32b_x32(lane_base + L::O_COL, o0);
    tmem_ld_32x32b_x32(lane_base + L::O_COL + 32, o1);
    tmem_ld_wait();
    tcgen05_fence_before();
    if (q0 + row < T) {
      const float inv = 1.f / sum;
      uint4* op = reinterpret_cast<uint4*>(out + ((size_t)b * T + q0 + row) * out_ld + head * DH);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_h<F16>(__uint_as_float(o0[8 * i + 0]) * inv, __uint_as_float(o0[8 * i + 1]) * inv);
        u.y = pack_h<F16>(__uint_as_float(o0[8 * i + 2]) * inv, __uint_as_float(o0[8 * i + 3]) * inv);
        u.z = pack_h<F16>(__uint_as_float(o0[8 * i + 4]) * inv, __uint_as_float(o0[8 * i + 5]) * inv);
        u.w = pack_h<F16>(__uint_as_float(o0[8 * i + 6]) * inv, __uint_as_float(o0[8 * i + 7]) * inv);
        op[i] = u;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_h<F16>(__uint_as_float(o1[8 * i + 0]) * inv, __uint_as_float(o1[8 * i + 1]) * inv);
        u.y = pack_h<F16>(__uint_as_float(o1[8 * i + 2]) * inv, __uint_as_float(o1[8 * i + 3]) * inv);
        u.z = pack_h<F16>(__uint_as_float(o1[8 * i + 4]) * inv, __uint_as_float(o1[8 * i + 5]) * inv);
        u.w = pack_h<F16>(__uint_as_float(o1[8 * i + 6]) * inv, __uint_as_float(o1[8 * i + 7]) * inv);
        op[4 + i] = u;
      }
    }
  }
  __syncthreads();
  if (warp == 4) { tcgen05_fence_after(); tmem_dealloc(tmem_base, L::TMEM_COLS); }
}

template <int T, bool F16>
void launch_oaca_t(const bf16* qkv, const bf16* pos, int pos_batch, const bf16* klay, const bf16* vlay, int n_layout,
                   const View& out, int B, int C, cudaStream_t s) {
  using L = OacaCfg<T>;
  static bool configured = false;
  if (!configured) {
    LIDM_CUDA_CHECK(cudaFuncSetAttribute(oaca_attention_kernel<T, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM_TOTAL));
    configured = true;
  }
  const int heads = C / DH;
  CUtensorMap tmQKV = make_tma_3d(qkv, 3 * C, T, B, (uint64_t)3 * C * 2, (uint64_t)T * 3 * C * 2, DH, 64, 128);
  CUtensorMap tmPos = make_tma_3d(pos, C, T, pos_batch, (uint64_t)C * 2, (uint64_t)T * C * 2, DH, 64, 128);
  CUtensorMap tmKL = make_tma_3d(klay, 2 * C, LAY, B, (uint64_t)2 * C * 2, (uint64_t)LAY * 2 * C * 2, DH, LAY, 128);
  CUtensorMap tmVL = make_tma_3d(vlay, C, LAY, B, (uint64_t)C * 2, (uint64_t)LAY * C * 2, DH, LAY, 128);
  dim3 grid((T + QROWS - 1) / QROWS, heads, B);
  oaca_attention_kernel<T, F16><<<grid, 160, L::SMEM_TOTAL, s>>>(tmQKV, tmPos, tmKL, tmVL, out.p, out.ld, C, pos_batch > 1 ? 1 : 0,
                                                                n_layout);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

// ------------------------------------------------------------------------------------------------- conditioning prep
// y[b, c, l] = GroupNorm32( bias[c] + sum_e W[c, e] src[b, e, l] ) * scale  ->  dst[(b * rows + l) * dst_ld + dst_col + c]
// One CTA per (sample, group): the group's (C / 32) x L projections live in shared memory between the two passes.
__global__ void oaca_pos_kernel(const float* __restrict__ src, int E, int Lt, const float* __restrict__ W, const float* __restrict__ bias,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int C, float eps, float scale,
                                bf16* __restrict__ dst, int rows, int dst_ld, int dst_col, bool f16) {
  extern __shared__ float sh[];                  // [cpg][Lt] projections, then 64 floats of reduction scratch
  const int b = blockIdx.y, g = blockIdx.x;
  const int cpg = C / 32;
  float* proj = sh;
  float* red = sh + cpg * Lt;
  const float* sb = src + (size_t)b * E * Lt;
  float s = 0.f, q = 0.f;
  for (int i = threadIdx.x; i < cpg * Lt; i += blockDim.x) {
    const int cl = i / Lt, l = i - cl * Lt;
    const int c = g * cpg + cl;
    const float* wr = W + (size_t)c * E;
    float acc = bias[c];
    for (int e = 0; e < E; ++e) acc = fmaf(__ldg(wr + e), __ldg(sb + (size_t)e * Lt + l), acc);
    proj[i] = acc;
    s += acc; q += acc * acc;
  }
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s; red[32 + (threadIdx.x >> 5)] = q; }
  __syncthreads();
  s = 0.f; q = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { s += red[i]; q += red[32 + i]; }
  const float n = (float)(cpg * Lt);
  const float mean = s / n;
  const float rstd = rsqrtf(fmaxf(q / n - mean * mean, 0.f) + eps);
  for (int i = threadIdx.x; i < cpg * Lt; i += blockDim.x) {
    const int cl = i % cpg, l = i / cpg;           // channel fastest: neighbouring threads write neighbouring channels
    const int c = g * cpg + cl;
    const float v = ((proj[cl * Lt + l] - mean) * rstd * gamma[c] + beta[c]) * scale;
    store_hr(dst + ((size_t)b * rows + l) * dst_ld + dst_col + c, v, f16);
  }
}

// layout tokens: content = (xf_out + GroupNorm32(obj_class_embedding)) / 2; [k | v] = Wc content + bc
//   k * scale -> klay[(b * 16 + l) * 2C + c],  v -> vlay[(b * 16 + l) * C + c]     (one CTA per sample)
__global__ void oaca_layout_kv_kernel(const float* __restrict__ xf_out, const float* __restrict__ cls, int E, int Lt,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                      const float* __restrict__ Wc, const float* __restrict__ bc, int C, float scale,
                                      bf16* __restrict__ klay, bf16* __restrict__ vlay, bool f16) {
  extern __shared__ float sh[];                  // content [E][Lt], then per-group (mean, rstd) [32][2]
  const int b = blockIdx.x;
  float* content = sh;
  float* stat = sh + E * Lt;
  const float* cb = cls + (size_t)b * E * Lt;
  const float* xb = xf_out + (size_t)b * E * Lt;
  const int cpg = E / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < 32; g += blockDim.x >> 5) {
    float s = 0.f, q = 0.f;
    for (int i = lane; i < cpg * Lt; i += 32) { const float v = cb[(size_t)g * cpg * Lt + i]; s += v; q += v * v; }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (lane == 0) {
      const float n = (float)(cpg * Lt), mean = s / n;
      stat[2 * g] = mean;
      stat[2 * g + 1] = rsqrtf(fmaxf(q / n - mean * mean, 0.f) + eps);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < E * Lt; i += blockDim.x) {
    const int e = i / Lt, g = e / cpg;
    content[i] = (xb[i] + ((cb[i] - stat[2 * g]) * stat[2 * g + 1] * gamma[e] + beta[e])) * 0.5f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C * Lt; i += blockDim.x) {
    const int n = i % (2 * C), l = i / (2 * C);
    const float* wr = Wc + (size_t)n * E;
    float acc = bc[n];
    for (int e = 0; e < E; ++e) acc = fmaf(__ldg(wr + e), content[e * Lt + l], acc);
    if (n < C) store_hr(klay + ((size_t)b * LAY + l) * 2 * C + n, acc * scale, f16);
    else store_hr(vlay + ((size_t)b * LAY + l) * C + (n - C), acc, f16);
  }
}

// F.avg_pool2d(kernel 2, stride 2) on a halo-free channels-last tensor; 16-byte vectors
__global__ void avgpool2_kernel(const bf16* __restrict__ x, int B, int H, int W, int xld, int C, bf16* __restrict__ y, int yld,
                                bool f16) {
  const int vec = C >> 3, Ho = H / 2, Wo = W / 2;
  const int64_t total = (int64_t)B * Ho * Wo * vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vec);
    int64_t r = i / vec;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const bf16* base = x + ((size_t)(b * H + 2 * ho) * W + 2 * wo) * xld;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(base) + cv), bq = __ldg(reinterpret_cast<const uint4*>(base + xld) + cv);
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(base + (size_t)W * xld) + cv),
                d = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(W + 1) * xld) + cv);
    const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {bq.x, bq.y, bq.z, bq.w}, uc[4] = {c.x, c.y, c.z, c.w},
                   ud[4] = {d.x, d.y, d.z, d.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = unpack_hr(ua[k], f16), fb = unpack_hr(ub[k], f16), fc = unpack_hr(uc[k], f16), fd = unpack_hr(ud[k], f16);
      o[k] = pack_hr(((fa.x + fb.x) + (fc.x + fd.x)) * 0.25f, ((fa.y + fb.y) + (fc.y + fd.y)) * 0.25f, f16);
    }
    reinterpret_cast<uint4*>(y + ((size_t)(b * Ho + ho) * Wo + wo) * yld)[cv] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------- R2DM FIR resampling
// Resample(down=2) (unets/ops.py:52-143, window [1,3,3,1]/8 per axis): y[ho][wo] = sum_ij k_i k_j x[2ho+i-1][(2wo+j-1) mod W]
// ---------------------------------------------------------------------------------------------- R2DM input convolution
// in_conv of EfficientUNet (efficient_unet.py:262-270): a ring-padded 3x3 convolution over [x (Cx image channels) | Fourier
// features of the polar coordinates (constant)].  The coordinate channels do not depend on the sample or the step, so their
// contribution plus the bias is a constant per-pixel map computed once (eff_in_map_kernel, fp32); a step only convolves the
// Cx = 2 image channels (18 MACs per output) on CUDA cores and adds the map - instead of a 9 x 34 -> 320-column im2col of the
// full-resolution input (1.3 GB written per evaluation at B = 32: 2.5 ms) and a GEMM over it.
__global__ void eff_in_map_kernel(const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ cenc, int Cx,
                                  int Ce, int H, int W, int C0, float* __restrict__ map) {
  const int64_t total = (int64_t)H * W * C0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % C0);
    const int pix = (int)(i / C0);
    const int hh = pix / W, ww = pix - hh * W;
    float acc = bias != nullptr ? bias[n] : 0.f;
    for (int c = 0; c < Ce; ++c) {
      const float* wc = w + ((size_t)n * (Cx + Ce) + Cx + c) * 9;
      const float* f = cenc + (size_t)c * H * W;
      for (int ky = 0; ky < 3; ++ky) {
        const int y = hh + ky - 1;
        if (y < 0 || y >= H) continue;                      // zero padding in elevation
        for (int kx = 0; kx < 3; ++kx) {
          int x = ww + kx - 1;
          x = x < 0 ? x + W : (x >= W ? x - W : x);         // ring padding in azimuth
          acc = fmaf(wc[ky * 3 + kx], f[(size_t)y * W + x], acc);
        }
      }
    }
    map[i] = acc;
  }
}

// one thread: 8 output channels of PX = 4 consecutive pixels of a row (the 8 x CX x 9 weights are read once per 4 pixels as
// float4 from shared memory, the 3 x 6 input window once); 8 threads cover the 64 channels of the pixel group
template <int CX>
__global__ void __launch_bounds__(256) eff_in_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ map,
                                                          int B, int H, int W, int C0, bf16* __restrict__ out, int out_ld, bool f16) {
  constexpr int PX = 4, KW = CX * 9;
  extern __shared__ float4 sw4[];                            // [C0][KW] image-channel weights (KW * 8 floats per thread group: 16-byte aligned)
  float* sw = reinterpret_cast<float*>(sw4);
  for (int i = threadIdx.x; i < C0 * KW; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int groups = C0 >> 3, wq = W / PX;
  const int64_t total = (int64_t)B * H * wq * groups;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    int64_t r = i / groups;
    const int w0 = (int)(r % wq) * PX;
    r /= wq;
    const int hh = (int)(r % H);
    const int b = (int)(r / H);
    float in[CX][3][PX + 2];
#pragma unroll
    for (int c = 0; c < CX; ++c) {
      const float* xc = x + ((size_t)b * CX + c) * H * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int y = hh + ky - 1;
        const bool ok = y >= 0 && y < H;                    // zero padding in elevation
#pragma unroll
        for (int kx = 0; kx < PX + 2; ++kx) {
          int xx = w0 + kx - 1;
          xx = xx < 0 ? xx + W : (xx >= W ? xx - W : xx);   // ring padding in azimuth
          in[c][ky][kx] = ok ? __ldg(xc + (size_t)y * W + xx) : 0.f;
        }
      }
    }
    float acc[PX][8];
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      const float4* m4 = reinterpret_cast<const float4*>(map + ((size_t)hh * W + w0 + px) * C0 + g * 8);
      const float4 ma = __ldg(m4), mb = __ldg(m4 + 1);
      acc[px][0] = ma.x; acc[px][1] = ma.y; acc[px][2] = ma.z; acc[px][3] = ma.w;
      acc[px][4] = mb.x; acc[px][5] = mb.y; acc[px][6] = mb.z; acc[px][7] = mb.w;
    }
    const float4* wg4 = sw4 + (size_t)g * 8 * KW / 4;       // 8 channels x KW weights, contiguous
#pragma unroll
    for (int q = 0; q < 8 * KW / 4; ++q) {
      const float4 wv = wg4[q];
      const float wk[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = q * 4 + e, j = idx / KW, k = idx % KW, c = k / 9, ky = (k % 9) / 3, kx = k % 3;
#pragma unroll
        for (int px = 0; px < PX; ++px) acc[px][j] = fmaf(wk[e], in[c][ky][px + kx], acc[px][j]);
      }
    }
#pragma unroll
    for (int px = 0; px < PX; ++px) {
      uint4 o;
      o.x = pack_hr(acc[px][0], acc[px][1], f16); o.y = pack_hr(acc[px][2], acc[px][3], f16);
      o.z = pack_hr(acc[px][4], acc[px][5], f16); o.w = pack_hr(acc[px][6], acc[px][7], f16);
      *reinterpret_cast<uint4*>(out + (((size_t)b * H + hh) * W + w0 + px) * out_ld + g * 8) = o;
    }
  }
}

__global__ void fir_down2_kernel(const bf16* __restrict__ x, int B, int H, int W, int xld, int C, bf16* __restrict__ y, int yld,
                                 bool f16) {
  const int vec = C >> 3, Ho = H / 2, Wo = W / 2;
  const int64_t total = (int64_t)B * Ho * Wo * vec;
  const float k[4] = {0.125f, 0.375f, 0.375f, 0.125f};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vec);
    int64_t r = i / vec;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int hs = 2 * ho + a - 1;
      if (hs < 0 || hs >= H) continue;
      float row[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int ws = 2 * wo + c - 1;
        ws = ws < 0 ? ws + W : (ws >= W ? ws - W : ws);
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(b * H + hs) * W + ws) * xld) + cv);
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_hr(uu[q], f16);
          row[2 * q] = fmaf(k[c], f.x, row[2 * q]);
          row[2 * q + 1] = fmaf(k[c], f.y, row[2 * q + 1]);
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = fmaf(k[a], row[q], acc[q]);
    }
    uint4 o;
    o.x = pack_hr(acc[0], acc[1], f16); o.y = pack_hr(acc[2], acc[3], f16);
    o.z = pack_hr(acc[4], acc[5], f16); o.w = pack_hr(acc[6], acc[7], f16);
    reinterpret_cast<uint4*>(y + ((size_t)(b * Ho + ho) * Wo + wo) * yld)[cv] = o;
  }
}

// Resample(up=2): zero insertion + [1,3,3,1]/4 per axis = y[2m] = x[m-1]/4 + 3 x[m]/4, y[2m+1] = 3 x[m]/4 + x[m+1]/4 on both axes
__global__ void fir_up2_kernel(const bf16* __restrict__ x, int B, int H, int W, int xld, int C, bf16* __restrict__ y, int yld,
                               int yhl, int yhr, bool f16) {
  const int vec = C >> 3, Ho = H * 2, Wo = W * 2;
  const int64_t total = (int64_t)B * Ho * Wo * vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vec);
    int64_t r = i / vec;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    // the two source rows / columns and their weights (the first listed is the far one, weight 1/4)
    const int h_far = (ho & 1) ? ho / 2 + 1 : ho / 2 - 1, h_near = ho / 2;
    int w_far = (wo & 1) ? wo / 2 + 1 : wo / 2 - 1;
    const int w_near = wo / 2;
    w_far = w_far < 0 ? w_far + W : (w_far >= W ? w_far - W : w_far);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int hs = a ? h_near : h_far;
      const float kh = a ? 0.75f : 0.25f;
      if (hs < 0 || hs >= H) continue;
      float row[8];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ws = c ? w_near : w_far;
        const float kw = c ? 0.75f : 0.25f;
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(b * H + hs) * W + ws) * xld) + cv);
        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 f = unpack_hr(uu[q], f16);
          row[2 * q] = c ? fmaf(kw, f.x, row[2 * q]) : kw * f.x;
          row[2 * q + 1] = c ? fmaf(kw, f.y, row[2 * q + 1]) : kw * f.y;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = fmaf(kh, row[q], acc[q]);
    }
    uint4 o;
    o.x = pack_hr(acc[0], acc[1], f16); o.y = pack_hr(acc[2], acc[3], f16);
    o.z = pack_hr(acc[4], acc[5], f16); o.w = pack_hr(acc[6], acc[7], f16);
    const size_t rowbase = (size_t)(b * Ho + ho) * (Wo + yhl + yhr);      // circular halo columns for the conv that follows
    reinterpret_cast<uint4*>(y + (rowbase + wo + yhl) * yld)[cv] = o;
    if (wo < yhr) reinterpret_cast<uint4*>(y + (rowbase + Wo + yhl + wo) * yld)[cv] = o;
    if (wo >= Wo - yhl) reinterpret_cast<uint4*>(y + (rowbase + (wo - (Wo - yhl))) * yld)[cv] = o;
  }
}

// ------------------------------------------------------------------------------------------------- layout encoder
// LayoutTransformerEncoder.forward (lidm/modules/encoders/layout_encoder.py:222-281) with its Transformer /
// ResidualAttentionBlock / QKVMultiheadAttention / MLP (:32-137) for the shipped condition types (obj_class, obj_bbox,
// is_valid_obj; no positional embedding, no key padding mask): 13 tokens of width H per sample - one CTA per sample keeps
// the token matrix in shared memory for the whole stack, fp32 throughout (runs once per conditioning).
// out[l][n] (+)= act(bias[n] + sum_k W[n][k] in[l][k]): one warp per output feature, lanes over k (coalesced weight rows)
template <int MODE>   // 0 store, 1 accumulate into out (residual), 2 store GELU (exact erf form)
__device__ __forceinline__ void enc_linear(const float* __restrict__ in, int ldi, int K, const float* __restrict__ W,
                                           const float* __restrict__ bias, int N, float* __restrict__ out, int ldo, int L) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int n = warp; n < N; n += nw) {
    float acc[LAY];
#pragma unroll
    for (int l = 0; l < LAY; ++l) acc[l] = 0.f;
    const float* wr = W + (size_t)n * K;
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int l = 0; l < LAY; ++l)
        if (l < L) acc[l] = fmaf(wv, in[l * ldi + k], acc[l]);
    }
#pragma unroll
    for (int l = 0; l < LAY; ++l) {
      float v = acc[l];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == l && l < L) {
        v += bias[n];
        if (MODE == 2) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
        if (MODE == 1) out[l * ldo + n] += v;
        else out[l * ldo + n] = v;
      }
    }
  }
}

__device__ __forceinline__ void enc_layernorm(const float* __restrict__ x, float* __restrict__ y, int H, int L,
                                              const float* __restrict__ g, const float* __restrict__ b) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int l = warp; l < L; l += nw) {
    float s = 0.f;
    for (int k = lane; k < H; k += 32) s += x[l * H + k];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)H;
    float q = 0.f;
    for (int k = lane; k < H; k += 32) { const float d = x[l * H + k] - mean; q += d * d; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)H + 1e-5f);
    for (int k = lane; k < H; k += 32) y[l * H + k] = (x[l * H + k] - mean) * rstd * g[k] + b[k];
  }
}

struct EncLayerP { const float *ln1_g, *ln1_b, *qkv_w, *qkv_b, *proj_w, *proj_b, *ln2_g, *ln2_b, *fc_w, *fc_b, *fc2_w, *fc2_b; };

__global__ void __launch_bounds__(512)
layout_encoder_kernel(const float* __restrict__ layout, int L, int H, int heads, int n_layers, const EncLayerP* __restrict__ layers,
                      const float* __restrict__ cls_emb, int n_classes, const float* __restrict__ be_w, const float* __restrict__ be_b,
                      const float* __restrict__ bx_w, const float* __restrict__ bx_b, const float* __restrict__ fln_g,
                      const float* __restrict__ fln_b, const float* __restrict__ tp_w, const float* __restrict__ tp_b, int out_dim,
                      float* __restrict__ xf_proj, float* __restrict__ xf_out, float* __restrict__ cls_out, float* __restrict__ bbox_out) {
  extern __shared__ float sh[];
  float* x = sh;                    // [L][H] token stream
  float* y = x + LAY * H;           // [L][H] LayerNorm output / attention output
  float* big = y + LAY * H;         // [L][4H] qkv (3H) or MLP hidden (4H)
  const int b = blockIdx.x;
  const float* lay = layout + (size_t)b * L * 13;
  // token embedding: class embedding + Linear(bbox_2d) + Linear(bbox) (layout_encoder.py:233-249)
  for (int i = threadIdx.x; i < L * H; i += blockDim.x) {
    const int l = i / H, k = i - l * H;
    const float* t = lay + l * 13;
    int c = (int)t[12];
    c = c < 0 ? 0 : (c >= n_classes ? n_classes - 1 : c);
    const float ce = cls_emb[(size_t)c * H + k];
    float e2 = be_b[k];
#pragma unroll
    for (int j = 0; j < 4; ++j) e2 = fmaf(be_w[k * 4 + j], t[8 + j], e2);
    float e8 = bx_b[k];
#pragma unroll
    for (int j = 0; j < 8; ++j) e8 = fmaf(bx_w[k * 8 + j], t[j], e8);
    x[i] = ce + e2 + e8;
    cls_out[((size_t)b * H + k) * L + l] = ce;
    bbox_out[((size_t)b * H + k) * L + l] = e2;
  }
  __syncthreads();
  const int ch = H / heads;
  const float scale2 = rsqrtf((float)ch);            // (ch^-1/4)^2
  for (int li = 0; li < n_layers; ++li) {
    const EncLayerP P = layers[li];
    enc_layernorm(x, y, H, L, P.ln1_g, P.ln1_b);
    __syncthreads();
    enc_linear<0>(y, H, H, P.qkv_w, P.qkv_b, 3 * H, big, 3 * H, L);
    __syncthreads();
    // QKVMultiheadAttention (layout_encoder.py:65-84): per head [q | k | v], softmax in fp32
    for (int i = threadIdx.x; i < heads * L; i += blockDim.x) {
      const int hd = i / L, t = i - hd * L;
      const float* q = big + t * 3 * H + hd * 3 * ch;
      float sc[LAY];
      float mx = -INFINITY;
      for (int s2 = 0; s2 < L; ++s2) {
        const float* k = big + s2 * 3 * H + hd * 3 * ch + ch;
        float a = 0.f;
        for (int c = 0; c < ch; ++c) a = fmaf(q[c], k[c], a);
        a *= scale2;
        sc[s2] = a;
        mx = fmaxf(mx, a);
      }
      float sum = 0.f;
      for (int s2 = 0; s2 < L; ++s2) { sc[s2] = expf(sc[s2] - mx); sum += sc[s2]; }
      const float inv = 1.f / sum;
      for (int c = 0; c < ch; ++c) {
        float a = 0.f;
        for (int s2 = 0; s2 < L; ++s2) a = fmaf(sc[s2] * inv, big[s2 * 3 * H + hd * 3 * ch + 2 * ch + c], a);
        y[t * H + hd * ch + c] = a;
      }
    }
    __syncthreads();
    enc_linear<1>(y, H, H, P.proj_w, P.proj_b, H, x, H, L);
    __syncthreads();
    enc_layernorm(x, y, H, L, P.ln2_g, P.ln2_b);
    __syncthreads();
    enc_linear<2>(y, H, H, P.fc_w, P.fc_b, 4 * H, big, 4 * H, L);
    __syncthreads();
    enc_linear<1>(big, 4 * H, 4 * H, P.fc2_w, P.fc2_b, H, x, H, L);
    __syncthreads();
  }
  if (fln_g != nullptr) {
    enc_layernorm(x, y, H, L, fln_g, fln_b);
    __syncthreads();
  } else {
    for (int i = threadIdx.x; i < L * H; i += blockDim.x) y[i] = x[i];
    __syncthreads();
  }
  for (int i = threadIdx.x; i < L * H; i += blockDim.x) {
    const int l = i / H, k = i - l * H;
    xf_out[((size_t)b * H + k) * L + l] = y[i];
  }
  enc_linear<0>(y, H, H, tp_w, tp_b, out_dim, xf_proj + (size_t)b * out_dim, out_dim, 1);   // transformer_proj(x[:, 0])
}

// image_patch_bbox_embedding_for_resolution{rows} = Linear_bbox_emb(patch boxes) (layout_encoder.py:198-204, 251-257):
// out (1, H, rows * cols), boxes (x0, y0, x1, y1) of the rows x cols patch grid in unit coordinates (float64 -> float32)
__global__ void patch_table_kernel(const float* __restrict__ be_w, const float* __restrict__ be_b, int H, int rows, int cols,
                                   float* __restrict__ out) {
  const int n = rows * cols;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * n; i += gridDim.x * blockDim.x) {
    const int k = i / n, pch = i - k * n;
    const int pi = pch / cols, pj = pch - pi * cols;
    const double di = 1.0 / rows, dj = 1.0 / cols;
    const float box[4] = {(float)(dj * pj), (float)(di * pi), (float)(dj * (pj + 1)), (float)(di * (pi + 1))};
    float e = be_b[k];
#pragma unroll
    for (int j = 0; j < 4; ++j) e = fmaf(be_w[k * 4 + j], box[j], e);
    out[i] = e;
  }
}

}  // namespace

void launch_layout_encoder(const float* layout, int B, int L, int H, int heads, int n_layers, const void* layers_dev,
                           const float* cls_emb, int n_classes, const float* be_w, const float* be_b, const float* bx_w,
                           const float* bx_b, const float* fln_g, const float* fln_b, const float* tp_w, const float* tp_b,
                           int out_dim, float* xf_proj, float* xf_out, float* cls_out, float* bbox_out, cudaStream_t s) {
  LIDM_REQUIRE(L >= 1 && L <= LAY && H % 32 == 0 && H % heads == 0, "layout encoder: 1..16 tokens, width a multiple of 32");
  const size_t sh = (size_t)LAY * H * 6 * sizeof(float);
  static size_t configured = 0;
  if (sh > configured) {
    LIDM_CUDA_CHECK(cudaFuncSetAttribute(layout_encoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    configured = sh;
  }
  layout_encoder_kernel<<<B, 512, sh, s>>>(layout, L, H, heads, n_layers, reinterpret_cast<const EncLayerP*>(layers_dev), cls_emb,
                                          n_classes, be_w, be_b, bx_w, bx_b, fln_g, fln_b, tp_w, tp_b, out_dim, xf_proj, xf_out,
                                          cls_out, bbox_out);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_patch_table(const float* be_w, const float* be_b, int H, int rows, int cols, float* out, cudaStream_t s) {
  const int total = H * rows * cols;
  patch_table_kernel<<<(total + 255) / 256, 256, 0, s>>>(be_w, be_b, H, rows, cols, out);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_oaca_attention(const bf16* qkv, const bf16* pos, int pos_batch, const bf16* klay, const bf16* vlay, int n_layout,
                           const View& out, int B, int T, int C, cudaStream_t s) {
  LIDM_REQUIRE(C % DH == 0 && out.C == C && out.B == B && out.H * out.W == T && out.hl == 0 && out.hr == 0 && out.ld % 8 == 0,
               "object-aware attention: output view");
  LIDM_REQUIRE(n_layout >= 1 && n_layout <= LAY, "object-aware attention: 1..16 layout tokens");
  LIDM_REQUIRE(pos_batch == 1 || pos_batch == B, "object-aware attention: positional tensor batch");
#define LIDM_OACA(TT)                                                                                   \
  do {                                                                                                  \
    if (out.f16) launch_oaca_t<TT, true>(qkv, pos, pos_batch, klay, vlay, n_layout, out, B, C, s);       \
    else launch_oaca_t<TT, false>(qkv, pos, pos_batch, klay, vlay, n_layout, out, B, C, s);              \
  } while (0)
  if (T == 256) LIDM_OACA(256);
  else if (T == 128) LIDM_OACA(128);
  else if (T == 64) LIDM_OACA(64);
  else throw Error(-1, "object-aware attention supports feature maps of 64, 128 or 256 positions (attention_ds of the "
                       "shipped layout2lidar configuration); got " + std::to_string(T));
#undef LIDM_OACA
}

void launch_oaca_pos(const float* src, int Bsrc, int E, int Lt, const float* W, const float* bias, const float* gamma,
                     const float* beta, int C, float scale, bf16* dst, int rows, int dst_ld, int dst_col, bool f16,
                     cudaStream_t s) {
  LIDM_REQUIRE(C % 32 == 0 && Lt >= 1, "positional projection shape");
  const size_t sh = ((size_t)(C / 32) * Lt + 64) * sizeof(float);
  LIDM_REQUIRE(sh <= 48 * 1024, "positional projection: group slab too large");
  oaca_pos_kernel<<<dim3(32, Bsrc), 256, sh, s>>>(src, E, Lt, W, bias, gamma, beta, C, 1e-5f, scale, dst, rows, dst_ld, dst_col, f16);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_oaca_layout_kv(const float* xf_out, const float* cls, int B, int E, int Lt, const float* gamma, const float* beta,
                           const float* Wc, const float* bc, int C, float scale, bf16* klay, bf16* vlay, bool f16,
                           cudaStream_t s) {
  LIDM_REQUIRE(E % 32 == 0 && Lt >= 1 && Lt <= LAY, "layout token projection shape");
  const size_t sh = ((size_t)E * Lt + 64) * sizeof(float);
  LIDM_REQUIRE(sh <= 48 * 1024, "layout token projection: shared memory");
  oaca_layout_kv_kernel<<<B, 256, sh, s>>>(xf_out, cls, E, Lt, gamma, beta, 1e-5f, Wc, bc, C, scale, klay, vlay, f16);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

static int fir_grid(int64_t total) {
  int64_t g = (total + 255) / 256;
  return (int)(g > 148 * 16 ? 148 * 16 : g);
}
void launch_eff_in_map(const float* w, const float* bias, const float* cenc, int Cx, int Ce, int H, int W, int C0, float* map,
                        cudaStream_t s) {
  eff_in_map_kernel<<<148 * 8, 256, 0, s>>>(w, bias, cenc, Cx, Ce, H, W, C0, map);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}
void launch_eff_in_conv(const float* x, const float* w, const float* map, int Cx, const View& out, cudaStream_t s) {
  LIDM_REQUIRE(out.hl + out.hr == 0 && out.wpitch == 0 && out.C % 8 == 0 && out.ld % 8 == 0, "R2DM input convolution: halo-free channels-last output");
  LIDM_REQUIRE(Cx == 2 || Cx == 1, "R2DM input convolution: one or two image channels");
  LIDM_REQUIRE(out.W % 4 == 0 && (8 * Cx * 9) % 4 == 0, "R2DM input convolution: W must be a multiple of 4");
  const int64_t total = (int64_t)out.B * out.H * (out.W / 4) * (out.C / 8);
  const int grid = fir_grid(total);
  const size_t sh = (size_t)out.C * Cx * 9 * sizeof(float);
  if (Cx == 2) eff_in_conv_kernel<2><<<grid, 256, sh, s>>>(x, w, map, out.B, out.H, out.W, out.C, out.p, out.ld, out.f16);
  else eff_in_conv_kernel<1><<<grid, 256, sh, s>>>(x, w, map, out.B, out.H, out.W, out.C, out.p, out.ld, out.f16);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}
void launch_fir_down2(const View& x, const View& y, cudaStream_t s) {
  LIDM_REQUIRE(x.hl + x.hr + y.hl + y.hr == 0 && y.H * 2 == x.H && y.W * 2 == x.W && y.C == x.C && y.B == x.B && x.C % 8 == 0 &&
                   x.f16 == y.f16 && x.wpitch == 0 && y.wpitch == 0, "FIR downsampling shapes");
  const int64_t total = (int64_t)y.B * y.H * y.W * (y.C / 8);
  fir_down2_kernel<<<fir_grid(total), 256, 0, s>>>(x.p, x.B, x.H, x.W, x.ld, x.C, y.p, y.ld, x.f16);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}
void launch_fir_up2(const View& x, const View& y, cudaStream_t s) {
  LIDM_REQUIRE(x.hl + x.hr == 0 && y.H == 2 * x.H && y.W == 2 * x.W && y.C == x.C && y.B == x.B && x.C % 8 == 0 && x.f16 == y.f16 &&
                   x.wpitch == 0 && y.wpitch == 0, "FIR upsampling shapes");
  const int64_t total = (int64_t)y.B * y.H * y.W * (y.C / 8);
  fir_up2_kernel<<<fir_grid(total), 256, 0, s>>>(x.p, x.B, x.H, x.W, x.ld, x.C, y.p, y.ld, y.hl, y.hr, x.f16);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_avgpool2(const View& x, const View& y, cudaStream_t s) {
  LIDM_REQUIRE(x.hl + x.hr + y.hl + y.hr == 0 && y.H * 2 == x.H && y.W * 2 == x.W && y.C == x.C && y.B == x.B && x.C % 8 == 0 &&
                   x.f16 == y.f16 && x.wpitch == 0 && y.wpitch == 0,
               "average pooling shapes");
  const int64_t total = (int64_t)y.B * y.H * y.W * (y.C / 8);
  int64_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  avgpool2_kernel<<<(int)g, 256, 0, s>>>(x.p, x.B, x.H, x.W, x.ld, x.C, y.p, y.ld, x.f16);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

}  // namespace lidm
