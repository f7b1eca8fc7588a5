// GroupNorm (+ optional SiLU) over channels-last bf16 activations, fp32 statistics.
// Replaces GroupNorm32 / normalization() + nn.SiLU (reference lidm/modules/basic.py:324-341, eps 1e-5) in the
// U-Net and Normalize() + nonlinearity() (reference lidm/modules/diffusion/model_lidm.py:35-41, eps 1e-6) in the
// decoder.  HBM-bound: pass 1 reads x once (per-CTA partial sums, combined in a fixed order => deterministic),
// pass 2 reads x once more and writes y (with its circular halo columns) once.  Works on concatenated views
// (ld > C), so the U-Net's skip `torch.cat` (openaimodel.py:745) is never materialised separately.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace lidm {

namespace {

// pixel index -> row: a shift when W is a power of two (every level of the shipped models), else a division.  The apply
// kernels split two pixel indices per 16-byte vector; at 64 channels that division was a third of their instructions.
__device__ __forceinline__ int row_of(int pix, int W) { return (W & (W - 1)) == 0 ? pix >> (31 - __clz(W)) : pix / W; }

// Each thread owns one 8-channel (16-byte) column `cv` of the tensor and walks pixels with a fixed stride.
// cpg = channels per group.  If cpg >= 8 the 8 channels fall in one group, otherwise in 8/cpg groups; NSUB == 8 is
// the generic path (one accumulator per channel).  All reductions run in a fixed order: results are bit-reproducible
// and, because the chunking depends only on (H*W, C), independent of the batch size.
template <int NSUB, bool F16>
__global__ void gn_stats_kernel(const bf16* __restrict__ x, int H, int W, int hl, int Wp, int ld, int C, int cpg,
                                int groups, int pix_per_cta, float* __restrict__ partials, int nchunks) {
  extern __shared__ float sh[];  // [blockDim][NSUB][2]
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int vec_per_pix = C >> 3;
  const int cv = threadIdx.x % vec_per_pix;
  const int prow = threadIdx.x / vec_per_pix;
  const int pstride = blockDim.x / vec_per_pix;
  float s[NSUB], q[NSUB];
#pragma unroll
  for (int i = 0; i < NSUB; ++i) { s[i] = 0.f; q[i] = 0.f; }
  const int HW = H * W;
  const int p0 = chunk * pix_per_cta;
  const int p1 = min(HW, p0 + pix_per_cta);
  auto accum = [&](const uint4& u) {
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_h<F16>(uu[i]);
      if constexpr (NSUB == 8) {
        s[2 * i] += f.x; q[2 * i] += f.x * f.x;
        s[2 * i + 1] += f.y; q[2 * i + 1] += f.y * f.y;
      } else {
        constexpr int per = 8 / NSUB;  // channels per sub-group
        const int sub = (2 * i) / per;
        s[sub] += f.x + f.y;
        q[sub] += f.x * f.x + f.y * f.y;
      }
    }
  };
  auto addr = [&](int pix) {
    const int h = row_of(pix, W), w = pix - h * W;
    return reinterpret_cast<const uint4*>(x + ((size_t)(b * H + h) * Wp + (w + hl)) * ld) + cv;
  };
  if (prow < pstride) {
    int pix = p0 + prow;
    for (; pix + 3 * pstride < p1; pix += 4 * pstride) {   // 4 independent 16-byte loads in flight
      const uint4 u0 = __ldg(addr(pix)), u1 = __ldg(addr(pix + pstride)), u2 = __ldg(addr(pix + 2 * pstride)),
                  u3 = __ldg(addr(pix + 3 * pstride));
      accum(u0); accum(u1); accum(u2); accum(u3);
    }
    for (; pix < p1; pix += pstride) accum(__ldg(addr(pix)));
  }
#pragma unroll
  for (int i = 0; i < NSUB; ++i) {
    sh[(threadIdx.x * NSUB + i) * 2 + 0] = s[i];
    sh[(threadIdx.x * NSUB + i) * 2 + 1] = q[i];
  }
  __syncthreads();
  // one owner thread per group sums its contributors in a fixed order
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float ts = 0.f, tq = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg;) {
      const int v = c >> 3;
      int slot, step;
      if (NSUB == 8) { slot = c & 7; step = 1; }
      else if (NSUB == 1) { slot = 0; step = 8; }
      else { slot = (c & 7) / (8 / NSUB); step = 8 / NSUB; }
      for (int pr = 0; pr < pstride; ++pr) {
        const int t = pr * vec_per_pix + v;
        ts += sh[(t * NSUB + slot) * 2 + 0];
        tq += sh[(t * NSUB + slot) * 2 + 1];
      }
      c += step;
    }
    float* out = partials + (((size_t)b * nchunks + chunk) * groups + g) * 2;
    out[0] = ts;
    out[1] = tq;
  }
}

// SiLU v * sigmoid(v) = h + h * tanh(h), h = v / 2: one MUFU op (tanh.approx, relative error 2^-11, below the bf16 rounding
// of the output) instead of the exponential + reciprocal pair - the apply kernels are memory-bound but spent a third of
// their issue on the MUFU pipe (16 results per clock per SM)
__device__ __forceinline__ float silu_f(float v) {
  const float h = 0.5f * v;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

template <bool F16>
__global__ void gn_apply_kernel(const bf16* __restrict__ x, int H, int W, int xhl, int xWp, int xld, bf16* __restrict__ y,
                                int yhl, int yhr, int yWp, int yld, int C, int cpg, int groups,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                                const float* __restrict__ partials, int nchunks, int pix_per_cta,
                                const float* __restrict__ film, int film_ld) {
  extern __shared__ float sh[];  // mean[groups], rstd[groups]
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.y;
  const int HW = H * W;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float s = 0.f, q = 0.f;
    const float* pp = partials + (size_t)b * nchunks * groups * 2 + g * 2;
    for (int c = 0; c < nchunks; ++c) { s += pp[(size_t)c * groups * 2]; q += pp[(size_t)c * groups * 2 + 1]; }
    const float n = (float)HW * (float)cpg;
    const float mean = s / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    sh[g] = mean;
    sh[groups + g] = rsqrtf(var + eps);
  }
  __syncthreads();
  const int vec_per_pix = C >> 3;
  const int cv = threadIdx.x % vec_per_pix;
  const int prow = threadIdx.x / vec_per_pix;
  const int pstride = blockDim.x / vec_per_pix;
  if (prow >= pstride) return;
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    const int g = c / cpg;
    float ga = __ldg(gamma + c) * sh[groups + g];
    float of = __ldg(beta + c) - sh[g] * ga;
    if (film != nullptr) {     // FiLM: norm(h) * (1 + scale) + shift (object_cross_unet.py:268-272)
      const float s1 = 1.f + __ldg(film + (size_t)b * film_ld + c);
      ga *= s1;
      of = fmaf(of, s1, __ldg(film + (size_t)b * film_ld + C + c));
    }
    sc[j] = ga;
    sf[j] = of;
  }
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(HW, p0 + pix_per_cta);
  auto addr = [&](int pix) {
    const int h = row_of(pix, W), w = pix - h * W;
    return reinterpret_cast<const uint4*>(x + ((size_t)(b * H + h) * xWp + (w + xhl)) * xld) + cv;
  };
  auto emit = [&](int pix, const uint4& u) {
    const int h = row_of(pix, W), w = pix - h * W;
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
    uint32_t oo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_h<F16>(uu[i]);
      float a = f.x * sc[2 * i] + sf[2 * i];
      float c = f.y * sc[2 * i + 1] + sf[2 * i + 1];
      if (silu) { a = silu_f(a); c = silu_f(c); }
      oo[i] = pack_h<F16>(a, c);
    }
    const uint4 o = make_uint4(oo[0], oo[1], oo[2], oo[3]);
    const size_t rowbase = (size_t)(b * H + h) * yWp;
    reinterpret_cast<uint4*>(y + (rowbase + w + yhl) * yld)[cv] = o;
    if (w < yhr) reinterpret_cast<uint4*>(y + (rowbase + W + yhl + w) * yld)[cv] = o;
    if (w >= W - yhl) reinterpret_cast<uint4*>(y + (rowbase + (w - (W - yhl))) * yld)[cv] = o;
  };
  int pix = p0 + prow;
  for (; pix + 3 * pstride < p1; pix += 4 * pstride) {   // 4 independent 16-byte loads in flight
    const uint4 u0 = __ldg(addr(pix)), u1 = __ldg(addr(pix + pstride)), u2 = __ldg(addr(pix + 2 * pstride)),
                u3 = __ldg(addr(pix + 3 * pstride));
    emit(pix, u0); emit(pix + pstride, u1); emit(pix + 2 * pstride, u2); emit(pix + 3 * pstride, u3);
  }
  for (; pix < p1; pix += pstride) emit(pix, __ldg(addr(pix)));
}

// One-pass GroupNorm for tensors whose per-(sample, group-chunk) slab fits in registers: one CTA owns `gpc` whole groups
// (gpc * cpg channels, vec = gpc * cpg / 8 sixteen-byte vectors per pixel) of one sample, every thread keeps its NV
// vectors in registers between the statistics and the normalisation, so x is read once and y written once (4 bytes per
// element instead of 6) in a single launch.  Reductions run in a fixed order (per-thread partials -> fixed lane
// assignment -> xor-shuffle tree): bit-reproducible and, one sample per CTA, independent of the batch size.
// gn_apply with the statistics reduced from the producers' granule partials (View::gst) instead of a stats pass.
constexpr int GN_GST_STAGE_MAX = 2048;   // float2 partials staged in shared memory (16 KB)

template <bool F16>
__global__ void gn_apply_gst_kernel(const bf16* __restrict__ x, int H, int W, int xhl, int xWp, int xld, bf16* __restrict__ y,
                                    int yhl, int yhr, int yWp, int yld, int C, int cpg, int groups,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int silu,
                                    const float* __restrict__ gst, int gst_ld, int gst_slots, int pix_per_cta,
                                    const float* __restrict__ film, int film_ld) {
  extern __shared__ float sh[];  // mean[groups], rstd[groups]
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.y;
  const int HW = H * W;
  const int gpg = cpg >> 3;      // granules per group
  const int vec_per_pix = C >> 3;
  const int cv = threadIdx.x % vec_per_pix;
  const int prow = threadIdx.x / vec_per_pix;
  const int pstride = blockDim.x / vec_per_pix;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(HW, p0 + pix_per_cta);
  auto addr = [&](int pix) {
    const int h = row_of(pix, W), w = pix - h * W;
    return reinterpret_cast<const uint4*>(x + ((size_t)(b * H + h) * xWp + (w + xhl)) * xld) + cv;
  };
  // the first batch of activations does not depend on the statistics: request it before the statistics prologue so
  // the two memory round trips overlap
  int pix = p0 + prow;
  const bool pre = prow < pstride && pix + 3 * pstride < p1;
  uint4 u0 = make_uint4(0, 0, 0, 0), u1 = u0, u2 = u0, u3 = u0;
  if (pre) {
    u0 = __ldg(addr(pix)); u1 = __ldg(addr(pix + pstride)); u2 = __ldg(addr(pix + 2 * pstride));
    u3 = __ldg(addr(pix + 3 * pstride));
  }
  // statistics prologue: the granule partials of this sample (gst_slots x groups x gpg pairs) are staged in shared memory by
  // ALL threads in one round of loads (32 threads walking 16 slots four loads at a time cost 3-4 dependent L2 round trips per
  // launch, a quarter of a small launch); one thread per group then sums them in the same slot-major order as before, so the
  // statistics keep their bits
  const int row2 = gst_ld >> 1;
  const int per_slot = groups * gpg;                 // float2 partials per slot
  const int n_part = gst_slots * per_slot;
  float2* stage = reinterpret_cast<float2*>(sh + 2 * groups);
  const bool staged = n_part <= GN_GST_STAGE_MAX;
  const float2* pb = reinterpret_cast<const float2*>(gst + (size_t)b * gst_slots * gst_ld);
  if (staged) {
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) {
      const int sl = i / per_slot, k = i - sl * per_slot;
      stage[i] = __ldg(pb + (size_t)sl * row2 + k);
    }
    __syncthreads();
  }
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float s = 0.f, q = 0.f;
    // order of the sums (unchanged since the first version, the statistics are part of the bit-reproducibility contract):
    // blocks of four slots, inside a block granule-major, then the slots of the block; leftover slots one by one
    auto part = [&](int sl, int k) {
      return staged ? stage[sl * per_slot + g * gpg + k] : __ldg(pb + (size_t)sl * row2 + (size_t)g * gpg + k);
    };
    int sl = 0;
    for (; sl + 4 <= gst_slots; sl += 4)
      for (int k = 0; k < gpg; ++k)
        for (int j = 0; j < 4; ++j) { const float2 a = part(sl + j, k); s += a.x; q += a.y; }
    for (; sl < gst_slots; ++sl)
      for (int k = 0; k < gpg; ++k) { const float2 a = part(sl, k); s += a.x; q += a.y; }
    const float n = (float)HW * (float)cpg;
    const float mean = s / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    sh[g] = mean;
    sh[groups + g] = rsqrtf(var + eps);
  }
  __syncthreads();
  if (prow >= pstride) return;
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    const int g = c / cpg;
    float ga = __ldg(gamma + c) * sh[groups + g];
    float of = __ldg(beta + c) - sh[g] * ga;
    if (film != nullptr) {     // FiLM: norm(h) * (1 + scale) + shift (object_cross_unet.py:268-272)
      const float s1 = 1.f + __ldg(film + (size_t)b * film_ld + c);
      ga *= s1;
      of = fmaf(of, s1, __ldg(film + (size_t)b * film_ld + C + c));
    }
    sc[j] = ga;
    sf[j] = of;
  }
  auto emit = [&](int pix, const uint4& u) {
    const int h = row_of(pix, W), w = pix - h * W;
    const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
    uint32_t oo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_h<F16>(uu[i]);
      float a = f.x * sc[2 * i] + sf[2 * i];
      float c = f.y * sc[2 * i + 1] + sf[2 * i + 1];
      if (silu) { a = silu_f(a); c = silu_f(c); }
      oo[i] = pack_h<F16>(a, c);
    }
    const uint4 o = make_uint4(oo[0], oo[1], oo[2], oo[3]);
    const size_t rowbase = (size_t)(b * H + h) * yWp;
    reinterpret_cast<uint4*>(y + (rowbase + w + yhl) * yld)[cv] = o;
    if (w < yhr) reinterpret_cast<uint4*>(y + (rowbase + W + yhl + w) * yld)[cv] = o;
    if (w >= W - yhl) reinterpret_cast<uint4*>(y + (rowbase + (w - (W - yhl))) * yld)[cv] = o;
  };
  if (pre) {
    emit(pix, u0); emit(pix + pstride, u1); emit(pix + 2 * pstride, u2); emit(pix + 3 * pstride, u3);
    pix += 4 * pstride;
  }
  // (requesting the next four pixels before normalising the current four - a register double buffer - was measured: 80
  // registers instead of 63 cost a resident CTA per SM and the C256 @16x128 apply went 35.1 -> 35.6 us)
  for (; pix + 3 * pstride < p1; pix += 4 * pstride) {
    u0 = __ldg(addr(pix)); u1 = __ldg(addr(pix + pstride)); u2 = __ldg(addr(pix + 2 * pstride));
    u3 = __ldg(addr(pix + 3 * pstride));
    emit(pix, u0); emit(pix + pstride, u1); emit(pix + 2 * pstride, u2); emit(pix + 3 * pstride, u3);
  }
  for (; pix < p1; pix += pstride) emit(pix, __ldg(addr(pix)));
}

template <int NV, bool F16>
__global__ void __launch_bounds__(512)
gn_fused_kernel(const bf16* __restrict__ x, int H, int W, int xhl, int xWp, int xld, bf16* __restrict__ y, int yhl, int yhr,
                int yWp, int yld, int cpg, int gpc, const float* __restrict__ gamma, const float* __restrict__ beta,
                float eps, int silu, const float* __restrict__ film, int film_ld, int C) {
  __shared__ float2 red[512];
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float2 stat[32];   // (mean, rstd) per local group
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * gpc * cpg;
  const int vec = (gpc * cpg) >> 3, vg = cpg >> 3;
  const int PR = blockDim.x / vec;                 // pixel rows walked in parallel
  const int nthr = PR * vec;
  const bool active = threadIdx.x < nthr;
  const int cv = threadIdx.x % vec, prow = threadIdx.x / vec;
  const int HW = H * W;
  auto xaddr = [&](int pix) {
    const int h = row_of(pix, W), w = pix - h * W;
    return reinterpret_cast<const uint4*>(x + ((size_t)(b * H + h) * xWp + (w + xhl)) * xld + c0) + cv;
  };
  uint4 v[NV];
  float sm = 0.f, sq = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int pix = prow + k * PR;
    v[k] = make_uint4(0, 0, 0, 0);
    if (active && pix < HW) v[k] = __ldg(xaddr(pix));
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const uint32_t uu[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_h<F16>(uu[i]);
      sm += f.x + f.y;
      sq += f.x * f.x + f.y * f.y;
    }
  }
  red[threadIdx.x] = make_float2(sm, sq);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int g = warp; g < gpc; g += nwarps) {
    float ts = 0.f, tq = 0.f;
    const int n_contrib = PR * vg;
    for (int i = lane; i < n_contrib; i += 32) {
      const int pr = i / vg, vi = i - pr * vg;
      const float2 r = red[pr * vec + g * vg + vi];
      ts += r.x; tq += r.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ts += __shfl_xor_sync(0xffffffffu, ts, o);
      tq += __shfl_xor_sync(0xffffffffu, tq, o);
    }
    if (lane == 0) {
      const float n = (float)HW * (float)cpg;
      const float mean = ts / n;
      const float var = fmaxf(tq / n - mean * mean, 0.f);
      stat[g] = make_float2(mean, rsqrtf(var + eps));
    }
  }
  __syncthreads();
  if (!active) return;
  float sc[8], sf[8];
  {
    const float2 st = stat[cv / vg];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + cv * 8 + j;
      float ga = __ldg(gamma + c) * st.y;
      float of = __ldg(beta + c) - st.x * ga;
      if (film != nullptr) {
        const float s1 = 1.f + __ldg(film + (size_t)b * film_ld + c);
        ga *= s1;
        of = fmaf(of, s1, __ldg(film + (size_t)b * film_ld + C + c));
      }
      sc[j] = ga;
      sf[j] = of;
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int pix = prow + k * PR;
    if (pix >= HW) break;
    const int h = row_of(pix, W), w = pix - h * W;
    const uint32_t uu[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
    uint32_t oo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = unpack_h<F16>(uu[i]);
      float a = f.x * sc[2 * i] + sf[2 * i];
      float c = f.y * sc[2 * i + 1] + sf[2 * i + 1];
      if (silu) { a = silu_f(a); c = silu_f(c); }
      oo[i] = pack_h<F16>(a, c);
    }
    const uint4 o = make_uint4(oo[0], oo[1], oo[2], oo[3]);
    const size_t rowbase = (size_t)(b * H + h) * yWp;
    uint4* yp = reinterpret_cast<uint4*>(y + (rowbase + w + yhl) * yld + c0) + cv;
    *yp = o;
    if (w < yhr) *(reinterpret_cast<uint4*>(y + (rowbase + W + yhl + w) * yld + c0) + cv) = o;
    if (w >= W - yhl) *(reinterpret_cast<uint4*>(y + (rowbase + (w - (W - yhl))) * yld + c0) + cv) = o;
  }
}

}  // namespace

void launch_groupnorm(const View& x, const View& y, const float* gamma, const float* beta, float eps, int groups,
                      bool silu, float* partials, cudaStream_t s, const float* film, int film_ld) {
  const int C = x.C;
  LIDM_REQUIRE(C % 8 == 0 && C % groups == 0, "C must be a multiple of 8 and of the group count");
  LIDM_REQUIRE(y.C == C && y.B == x.B && y.H == x.H && y.W == x.W && x.f16 == y.f16, "GroupNorm in/out shape mismatch");
  LIDM_REQUIRE(x.ld % 8 == 0 && y.ld % 8 == 0, "ld alignment");
  const int cpg = C / groups;
  const bool regular = cpg >= 8 ? (cpg % 8 == 0) : (8 % cpg == 0 && cpg >= 2);
  const int vec = C / 8;
  int threads = (256 % vec == 0) ? 256 : ((384 % vec == 0) ? 384 : 0);
  if (threads == 0) { LIDM_REQUIRE(vec <= 1024, "C too large"); threads = vec; }
  const int HW = x.H * x.W;
  // one-pass register-resident kernel when a group chunk of one sample fits (the U-Net's tensors): pick the widest chunk
  // (most contiguous bytes per pixel) that keeps <= 16 vectors per thread and still fills the GPU
  static const bool two_pass = getenv("LIDM_GN_TWO_PASS") != nullptr;
  // (measured on B200, B=64: wins 20-25 % on the 4x32 level, loses on larger maps where the two-pass kernels' second
  // read hits L2 and their higher occupancy matters more than the saved pass)
  if (!two_pass && cpg % 8 == 0 && (cpg & (cpg - 1)) == 0 && HW <= 128) {
    const int vg = cpg / 8;
    int best = 0;
    for (int gpc = 1; gpc <= groups; gpc *= 2) {
      const int v = gpc * vg;
      if (v > 512) break;
      const int pr = 512 / v;
      const int nv = (HW + pr - 1) / pr;
      if (nv > 16) break;
      if (gpc > 1 && (long)x.B * (groups / gpc) < 296 && gpc * cpg * 2 > 64) break;   // keep >= 2 CTAs per SM once rows are >= 64 B
      best = gpc;
    }
    if (best > 0) {
      const int v = best * vg, pr = 512 / v, nv = (HW + pr - 1) / pr;
      dim3 grid(groups / best, x.B);
#define GN_FUSED(NV)                                                                                                  \
  do {                                                                                                                \
    if (x.f16) launch_pdl(gn_fused_kernel<NV, true>, grid, dim3(512), 0, s, x.p, x.H, x.W, x.hl, x.Wp(), x.ld, y.p, y.hl, y.hr, y.Wp(), y.ld, cpg, best, gamma, beta, eps, silu ? 1 : 0, film, film_ld, C); \
    else launch_pdl(gn_fused_kernel<NV, false>, grid, dim3(512), 0, s, x.p, x.H, x.W, x.hl, x.Wp(), x.ld, y.p, y.hl, y.hr, y.Wp(), y.ld, cpg, best, gamma, beta, eps, silu ? 1 : 0, film, film_ld, C); \
  } while (0)
      if (nv <= 4) GN_FUSED(4);
      else if (nv <= 8) GN_FUSED(8);
      else GN_FUSED(16);
#undef GN_FUSED
      LIDM_CUDA_CHECK(cudaGetLastError());
      LIDM_COUNT_LAUNCH(1);
      return;
    }
  }
  // chunking depends on the tensor shape only (never on B): results are batch-invariant
  const int pstride = threads / vec;
  int pix_per_cta = HW <= 512 ? 32 : 64;
  if (HW / pix_per_cta > GN_MAX_CHUNKS) pix_per_cta = (HW + GN_MAX_CHUNKS - 1) / GN_MAX_CHUNKS;
  if (pix_per_cta < pstride) pix_per_cta = pstride;
  const int nchunks = (HW + pix_per_cta - 1) / pix_per_cta;
  LIDM_REQUIRE(nchunks <= GN_MAX_CHUNKS, "GroupNorm chunking");
  dim3 grid(nchunks, x.B);
  const int nsub = !regular ? 8 : (cpg >= 8 ? 1 : 8 / cpg);
  const size_t shstats = (size_t)threads * nsub * 2 * sizeof(float);
  const size_t shbytes = groups * 2 * sizeof(float);
#define GN_STATS(NS, F)                                                                                              \
  launch_pdl(gn_stats_kernel<NS, F>, grid, dim3(threads), shstats, s, x.p, x.H, x.W, x.hl, x.Wp(), x.ld, C, cpg, groups, \
             pix_per_cta, partials, nchunks)
#define GN_STATS_F(F)                                                                                                 \
  do {                                                                                                                \
    if (nsub == 8) GN_STATS(8, F);                                                                                    \
    else if (nsub == 1) GN_STATS(1, F);                                                                               \
    else if (nsub == 2) GN_STATS(2, F);                                                                               \
    else GN_STATS(4, F);                                                                                              \
  } while (0)
  if (x.f16) GN_STATS_F(true);
  else GN_STATS_F(false);
#undef GN_STATS_F
#undef GN_STATS
  LIDM_CUDA_CHECK(cudaGetLastError());
  if (x.f16)
    launch_pdl(gn_apply_kernel<true>, grid, dim3(threads), shbytes, s, x.p, x.H, x.W, x.hl, x.Wp(), x.ld, y.p, y.hl, y.hr,
               y.Wp(), y.ld, C, cpg, groups, gamma, beta, eps, silu ? 1 : 0, (const float*)partials, nchunks, pix_per_cta, film,
               film_ld);
  else
    launch_pdl(gn_apply_kernel<false>, grid, dim3(threads), shbytes, s, x.p, x.H, x.W, x.hl, x.Wp(), x.ld, y.p, y.hl, y.hr,
               y.Wp(), y.ld, C, cpg, groups, gamma, beta, eps, silu ? 1 : 0, (const float*)partials, nchunks, pix_per_cta, film,
               film_ld);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(2);
}

void launch_groupnorm_from_gstats(const View& x, const View& y, const float* gamma, const float* beta, float eps, int groups,
                                  bool silu, cudaStream_t s, const float* film, int film_ld) {
  const int C = x.C;
  LIDM_REQUIRE(x.gst != nullptr && C % groups == 0 && (C / groups) % 8 == 0, "granule statistics need 8 | channels per group");
  LIDM_REQUIRE(y.C == C && y.B == x.B && y.H == x.H && y.W == x.W && x.ld % 8 == 0 && y.ld % 8 == 0, "GroupNorm shapes");
  const int cpg = C / groups;
  const int vec = C / 8;
  int threads = (256 % vec == 0) ? 256 : ((384 % vec == 0) ? 384 : 0);
  if (threads == 0) { LIDM_REQUIRE(vec <= 1024, "C too large"); threads = vec; }
  const int HW = x.H * x.W;
  const int pstride = threads / vec;
  // no reduction in this kernel, so the split is free to depend on the batch: aim at one resident wave of CTAs
  // (the kernel's registers, not the thread count, limit residency: ask the runtime rather than assume 2048 threads/SM -
  // a grid of 1.7 waves loses a quarter of the time to its ragged second wave)
  static int occ_cache[33] = {0};
  int& occ = occ_cache[threads / 32 <= 32 ? threads / 32 : 0];
  if (occ == 0) {
    LIDM_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_apply_gst_kernel<false>, threads,
                                                                  groups * 2 * sizeof(float) + GN_GST_STAGE_MAX * sizeof(float2)));
    if (occ < 1) occ = 1;
  }
  const int resident = 148 * occ;
  int nchunks = resident / x.B;
  if (nchunks < 1) nchunks = 1;
  if (nchunks > HW / pstride) nchunks = HW / pstride > 0 ? HW / pstride : 1;
  int pix_per_cta = (HW + nchunks - 1) / nchunks;
  pix_per_cta = (pix_per_cta + pstride - 1) / pstride * pstride;
  nchunks = (HW + pix_per_cta - 1) / pix_per_cta;
  dim3 grid(nchunks, x.B);
  LIDM_REQUIRE(x.f16 == y.f16, "GroupNorm element formats");
#define GN_GST(F)                                                                                                     \
  launch_pdl(gn_apply_gst_kernel<F>, grid, dim3(threads), groups * 2 * sizeof(float) + GN_GST_STAGE_MAX * sizeof(float2), s, x.p, x.H, x.W, x.hl, x.Wp(), x.ld, y.p, \
             y.hl, y.hr, y.Wp(), y.ld, C, cpg, groups, gamma, beta, eps, silu ? 1 : 0, (const float*)x.gst, x.gst_ld,       \
             x.gst_slots, pix_per_cta, film, film_ld)
  if (x.f16) GN_GST(true);
  else GN_GST(false);
#undef GN_GST
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

}  // namespace lidm
