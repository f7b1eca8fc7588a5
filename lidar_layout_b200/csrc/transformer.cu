// Memory-bound pieces of the SpatialTransformer path (reference lidm/modules/attention.py:196-215; the GEGLU gate is
// fused into the producing GEMM's epilogue, gemm_conv.cu) and the
// classifier-free-guidance update (reference lidm/models/diffusion/ddim.py:173-206).  Channels-last bf16 activations,
// fp32 arithmetic.  HBM-bound: every kernel reads its input once and writes its output once.
#include "common.h"
#include "ddim_math.cuh"
#include "ptx.cuh"

namespace lidm {

namespace {

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
inline int grid_for(int64_t work, int threads) {
  int64_t g = (work + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// One warp per token; each lane owns VEC 8-channel (16 B) vectors: channels [(lane + 32 v) * 8, +8).
// Two-pass statistics on registers (mean, then centred sum of squares) like ATen's LayerNorm kernel.
template <int VEC>
__global__ void layernorm_kernel(const bf16* __restrict__ x, int64_t xstride_tokens_unused, int H, int W, int xhl, int xWp,
                                 int xld, bf16* __restrict__ y, int yhl, int yWp, int yld, int C,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                 int64_t tokens) {
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  for (int64_t tok = (int64_t)blockIdx.x * warps_per_cta + (threadIdx.x >> 5); tok < tokens;
       tok += (int64_t)gridDim.x * warps_per_cta) {
    const int w = (int)(tok % W);
    const int64_t bh = tok / W;   // b * H + h
    const bf16* xp = x + ((size_t)bh * xWp + (w + xhl)) * xld;
    bf16* yp = y + ((size_t)bh * yWp + (w + yhl)) * yld;
    float v[VEC][8];
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int cv = lane + 32 * k;
      if (cv < nvec) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(xp) + cv);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        v[k][0] = a.x; v[k][1] = a.y; v[k][2] = b.x; v[k][3] = b.y; v[k][4] = c.x; v[k][5] = c.y; v[k][6] = d.x; v[k][7] = d.y;
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += v[k][i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[k][i] = 0.f;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      if (lane + 32 * k < nvec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[k][i] - mean; sq += d * d; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)C + eps);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int cv = lane + 32 * k;
      if (cv < nvec) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * cv), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * cv + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * cv), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * cv + 1);
        uint4 u;
        u.x = pack_bf16((v[k][0] - mean) * rstd * g0.x + b0.x, (v[k][1] - mean) * rstd * g0.y + b0.y);
        u.y = pack_bf16((v[k][2] - mean) * rstd * g0.z + b0.z, (v[k][3] - mean) * rstd * g0.w + b0.w);
        u.z = pack_bf16((v[k][4] - mean) * rstd * g1.x + b1.x, (v[k][5] - mean) * rstd * g1.y + b1.y);
        u.w = pack_bf16((v[k][6] - mean) * rstd * g1.z + b1.z, (v[k][7] - mean) * rstd * g1.w + b1.w);
        reinterpret_cast<uint4*>(yp)[cv] = u;
      }
    }
  }
}

__global__ void f32_rows_to_bf16_kernel(const float* __restrict__ x, int64_t n_valid, int64_t n_total, bf16* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __float2bfloat16(i < n_valid ? x[i] : 0.f);
}

__global__ void cfg_ddim_step_kernel(const float* __restrict__ x, const float* __restrict__ eps2, float scale,
                                     const float* __restrict__ noise, const float* __restrict__ coef_dev,
                                     float* __restrict__ x_prev, float* __restrict__ pred_x0, float* __restrict__ eps_out,
                                     int64_t n) {
  float coef[5] = {1.f, 1.f, 0.f, 0.f, 1.f};
  if (coef_dev != nullptr) {
#pragma unroll
    for (int i = 0; i < 5; ++i) coef[i] = __ldg(coef_dev + i);
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float eu = eps2[i], ec = eps2[n + i];
    // e_t = e_t_uncond + unconditional_guidance_scale * (e_t - e_t_uncond), rounded op by op like torch
    const float e = __fadd_rn(eu, __fmul_rn(scale, __fsub_rn(ec, eu)));
    if (eps_out != nullptr) eps_out[i] = e;
    if (x_prev != nullptr) {
      float xp, x0;
      ddim_update(x[i], e, noise != nullptr ? noise[i] : 0.f, coef, xp, x0);
      x_prev[i] = xp;
      if (pred_x0 != nullptr) pred_x0[i] = x0;
    }
  }
}

}  // namespace

void launch_layernorm(const View& x, const View& y, const float* gamma, const float* beta, float eps, cudaStream_t s) {
  const int C = x.C;
  LIDM_REQUIRE(C % 8 == 0 && C <= 2048, "LayerNorm: C must be a multiple of 8 and <= 2048");
  LIDM_REQUIRE(y.C == C && y.B == x.B && y.H == x.H && y.W == x.W && x.ld % 8 == 0 && y.ld % 8 == 0, "LayerNorm shapes");
  LIDM_REQUIRE(y.hl == 0 && y.hr == 0, "LayerNorm output has no halo");
  const int64_t tokens = (int64_t)x.B * x.H * x.W;
  const int threads = 256;
  const int grid = grid_for(tokens * 32, threads);
  const int vec = (C / 8 + 31) / 32;
#define LN_LAUNCH(V)                                                                                                 \
  layernorm_kernel<V><<<grid, threads, 0, s>>>(x.p, 0, x.H, x.W, x.hl, x.Wp(), x.ld, y.p, y.hl, y.Wp(), y.ld, C, gamma, \
                                               beta, eps, tokens)
  if (vec <= 1) LN_LAUNCH(1);
  else if (vec <= 2) LN_LAUNCH(2);
  else if (vec <= 4) LN_LAUNCH(4);
  else LN_LAUNCH(8);
#undef LN_LAUNCH
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_f32_rows_to_bf16(const float* x, int64_t rows, int64_t rows_pad, int cols, bf16* y, cudaStream_t s) {
  LIDM_REQUIRE(rows_pad >= rows && cols > 0, "f32_rows_to_bf16 shapes");
  f32_rows_to_bf16_kernel<<<grid_for(rows_pad * cols, 256), 256, 0, s>>>(x, rows * cols, rows_pad * cols, y);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_cfg_ddim_step(const float* x, const float* eps2, float scale, const float* noise, const float* coef_dev,
                          float* x_prev, float* pred_x0, float* eps_out, int64_t n, cudaStream_t s) {
  LIDM_REQUIRE(eps2 != nullptr && n > 0 && (x_prev == nullptr || (x != nullptr && coef_dev != nullptr)), "cfg step arguments");
  cfg_ddim_step_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, eps2, scale, noise, coef_dev, x_prev, pred_x0, eps_out, n);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

}  // namespace lidm
