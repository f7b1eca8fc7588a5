// HBM-bound elementwise / gather kernels of the sampling path (sm_100a):
//   ddim_step        DDIMSampler.p_sample_ddim update        (reference lidm/models/diffusion/ddim.py:191-206)
//   backproject      range2xyz / range2pcd geometry          (reference lidm/utils/lidar_utils.py:134-204,
//                                                              scripts/sample.py:29-35)
//   im2col_*         A-operand gathers for strided / 8-channel convs (CircularConv2d padding, basic.py:52-59)
//   upsample_*       F.interpolate nearest x2 (openaimodel.py:108-118) / bilinear align_corners (model_lidm.py:57-61)
//   vq               taming VectorQuantizer2 argmin + post_quant_conv (lidm/models/ae/vq.py:71-79, autoencoder.py:293-296)
//   time_embed       timestep_embedding + time_embed MLP + emb_layers (basic.py:278-296, openaimodel.py:509-514,262)
//   softmax_rows, weight packing, layout conversions.
#include <algorithm>

#include "common.h"
#include "ddim_math.cuh"
#include "ptx.cuh"

namespace lidm {

namespace {

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------------------------------ DDIM
__global__ void ddim_step_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                 const float* __restrict__ noise, const float* __restrict__ coef_dev,
                                 float* __restrict__ x_prev, float* __restrict__ pred_x0, int64_t n) {
  float coef[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) coef[i] = __ldg(coef_dev + i);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 xv = *reinterpret_cast<const float4*>(x + i);
      const float4 ev = *reinterpret_cast<const float4*>(eps + i);
      float4 nv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (noise != nullptr) nv = *reinterpret_cast<const float4*>(noise + i);
      float4 xp, x0;
      ddim_update(xv.x, ev.x, nv.x, coef, xp.x, x0.x);
      ddim_update(xv.y, ev.y, nv.y, coef, xp.y, x0.y);
      ddim_update(xv.z, ev.z, nv.z, coef, xp.z, x0.z);
      ddim_update(xv.w, ev.w, nv.w, coef, xp.w, x0.w);
      *reinterpret_cast<float4*>(x_prev + i) = xp;
      if (pred_x0 != nullptr) *reinterpret_cast<float4*>(pred_x0 + i) = x0;
    } else {
      for (int64_t k = i; k < n; ++k) {
        float xp, x0;
        ddim_update(x[k], eps[k], noise != nullptr ? noise[k] : 0.f, coef, xp, x0);
        x_prev[k] = xp;
        if (pred_x0 != nullptr) pred_x0[k] = x0;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ ancestral DDPM step
// p_sample (reference lidm/models/diffusion/ddpm.py:1090-1119) after the U-Net: per sample b with coefficients
// coef[b] = (sqrt_recip_ac, sqrt_recipm1_ac, posterior_mean_coef1, posterior_mean_coef2, mask * exp(0.5 logvar))
//   x0  = a x - b eps                 (predict_start_from_noise, ddpm.py:219-223; optional clamp to [-1, 1])
//   out = (c1 x0 + c2 x) + m noise    (q_posterior mean, ddpm.py:225-232, plus the masked noise term)
// every product and sum rounded on its own, in the reference's order: bit-identical to the eager tensor expression.
__global__ void ddpm_step_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                                 const float* __restrict__ coef, int64_t n_per_sample, int clip, float* __restrict__ x_prev,
                                 float* __restrict__ x_recon) {
  const int b = blockIdx.y;
  const float a = __ldg(coef + b * 5), bb = __ldg(coef + b * 5 + 1), c1 = __ldg(coef + b * 5 + 2), c2 = __ldg(coef + b * 5 + 3),
              m = __ldg(coef + b * 5 + 4);
  const size_t base = (size_t)b * n_per_sample;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_per_sample; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[base + i];
    float x0 = __fsub_rn(__fmul_rn(a, xv), __fmul_rn(bb, eps[base + i]));
    if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xv));
    x_prev[base + i] = __fadd_rn(mean, __fmul_rn(m, noise[base + i]));
    if (x_recon != nullptr) x_recon[base + i] = x0;
  }
}

// ------------------------------------------------------------------------------------------ back-projection
// img: (B,H,W) fp32 in [-1,1] (values outside are clipped, scripts/sample.py:31).  xyz: (B,3,H,W) fp32, -1 where
// masked (range2xyz semantics); mask: (B,H,W) uint8 = depth_min < d < depth_max (range2pcd keeps exactly these).
// Angles are evaluated in fp64 per row / column like the reference (np.float64 meshgrid), depth in fp32 like the
// reference (np.exp2 on a float32 array stays float32).
__global__ void backproject_kernel(const float* __restrict__ img, int H, int W, double fov_up, double fov_down,
                                   float dmin, float dmax, float depth_scale, int log_scale, int input_is_unit,
                                   float* __restrict__ xyz, uint8_t* __restrict__ mask) {
  extern __shared__ float trig[];  // cos(yaw)[W], sin(yaw)[W]
  const int b = blockIdx.z;
  const int h = blockIdx.y;
  const double fov_range = fabs(fov_down) + fabs(fov_up);
  const double kPi = 3.14159265358979323846;
  const double pitch = (1.0 - (double)h / (double)H) * fov_range - fabs(fov_down);
  const double cp = cos(pitch), sp = sin(pitch);
  const int w0 = blockIdx.x * blockDim.x * 4;
  for (int k = threadIdx.x; k < blockDim.x * 4; k += blockDim.x) {
    const int w = w0 + k;
    if (w < W) {
      const double yaw = kPi * (((double)w / (double)W) * 2.0 - 1.0);
      trig[k] = (float)(cos(yaw) * cp);
      trig[blockDim.x * 4 + k] = (float)(-sin(yaw) * cp);
    }
  }
  __syncthreads();
  const int k = threadIdx.x * 4;
  const int w = w0 + k;
  if (w >= W) return;
  const size_t HW = (size_t)H * W;
  const size_t base = (size_t)b * HW + (size_t)h * W + w;
  float v[4];
  if (w + 4 <= W) {
    const float4 t = *reinterpret_cast<const float4*>(img + base);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    for (int i = 0; i < 4; ++i) v[i] = (w + i < W) ? img[base + i] : 0.f;
  }
  float ox[4], oy[4], oz[4];
  uint8_t mk[4];
  const float spf = (float)sp;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float r = v[i];
    if (!input_is_unit) {
      r = fminf(fmaxf(r, -1.f), 1.f);
      r = __fdiv_rn(__fadd_rn(r, 1.f), 2.f);
    }
    float d = __fmul_rn(r, depth_scale);
    if (log_scale) d = __fsub_rn(exp2f(d), 1.f);
    const bool valid = (d > dmin) && (d < dmax);
    mk[i] = valid ? 1 : 0;
    ox[i] = valid ? trig[k + i] * d : -1.f;
    oy[i] = valid ? trig[blockDim.x * 4 + k + i] * d : -1.f;
    oz[i] = valid ? spf * d : -1.f;
  }
  float* o = xyz + (size_t)b * 3 * HW + (size_t)h * W + w;
  if (w + 4 <= W) {
    *reinterpret_cast<float4*>(o) = make_float4(ox[0], ox[1], ox[2], ox[3]);
    *reinterpret_cast<float4*>(o + HW) = make_float4(oy[0], oy[1], oy[2], oy[3]);
    *reinterpret_cast<float4*>(o + 2 * HW) = make_float4(oz[0], oz[1], oz[2], oz[3]);
    if (mask != nullptr) *reinterpret_cast<uchar4*>(mask + base) = make_uchar4(mk[0], mk[1], mk[2], mk[3]);
  } else {
    for (int i = 0; i < 4 && w + i < W; ++i) {
      o[i] = ox[i]; o[HW + i] = oy[i]; o[2 * HW + i] = oz[i];
      if (mask != nullptr) mask[base + i] = mk[i];
    }
  }
}

// ------------------------------------------------------------------------------------------ im2col (8-channel fp32 NCHW input)
// out[(b*H*W + h*W + w) * kpad + (ky*kw + kx)*C + c] = x[b][c][h+ky-pt][(w+kx-pl) mod W]  (0 outside H)
__global__ void im2col_nchw_f32_kernel(const float* __restrict__ x, int B, int C, int H, int W, int kh, int kw, int pl,
                                       int pt, bf16* __restrict__ out, int kpad, bool f16, bool zero_w) {
  // one thread = 8 consecutive k of one pixel (a 16-byte store); kpad is a multiple of 8
  const int kv = kpad >> 3;
  const int64_t total = (int64_t)B * H * W * kv;
  const int K = kh * kw * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k0 = (int)(i % kv) * 8;
    const int pixg = (int)(i / kv);
    const int w = pixg % W;
    const int h = (pixg / W) % H;
    const int b = pixg / (W * H);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      v[j] = 0.f;
      if (k < K) {
        const int tap = k / C, c = k - tap * C;
        const int ky = tap / kw, kx = tap - ky * kw;
        const int hs = h + ky - pt;
        int ws = w + kx - pl;
        const bool inside = ws >= 0 && ws < W;
        ws = ws < 0 ? ws + W : (ws >= W ? ws - W : ws);
        if (hs >= 0 && hs < H && (inside || !zero_w)) v[j] = __ldg(x + (((size_t)b * C + c) * H + hs) * W + ws);
      }
    }
    uint4 o;
    o.x = pack_hr(v[0], v[1], f16); o.y = pack_hr(v[2], v[3], f16); o.z = pack_hr(v[4], v[5], f16);
    o.w = pack_hr(v[6], v[7], f16);
    reinterpret_cast<uint4*>(out)[i] = o;
  }
}

// generic channels-last im2col (strided convs): out[(b,ho,wo)][tap*C + c]; 16-byte vectors
__global__ void im2col_nhwc_kernel(const bf16* __restrict__ x, int B, int H, int W, int hl, int Wp, int ld, int C,
                                   int kh, int kw, int sh, int sw, int pl, int pt, int Ho, int Wo, bf16* __restrict__ out) {
  const int vec = C >> 3;
  const int taps = kh * kw;
  const int64_t total = (int64_t)B * Ho * Wo * taps * vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vec);
    int64_t r = i / vec;
    const int tap = (int)(r % taps);
    r /= taps;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const int ky = tap / kw, kx = tap - ky * kw;
    const int hs = ho * sh + ky - pt;
    int ws = (wo * sw + kx - pl) % W;
    if (ws < 0) ws += W;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (hs >= 0 && hs < H) v = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(b * H + hs) * Wp + (ws + hl)) * ld) + cv);
    reinterpret_cast<uint4*>(out)[i] = v;
  }
}

// ------------------------------------------------------------------------------------------ resampling
__device__ __forceinline__ void store_with_halo(bf16* y, int b, int H, int W, int hl, int hr, int Wp, int ld, int h,
                                                int w, int cv, uint4 v) {
  const size_t rowbase = (size_t)(b * H + h) * Wp;
  reinterpret_cast<uint4*>(y + (rowbase + w + hl) * ld)[cv] = v;
  if (w < hr) reinterpret_cast<uint4*>(y + (rowbase + W + hl + w) * ld)[cv] = v;
  if (w >= W - hl) reinterpret_cast<uint4*>(y + (rowbase + (w - (W - hl))) * ld)[cv] = v;
}

__global__ void upsample_nearest_kernel(const bf16* __restrict__ x, int B, int H, int W, int xhl, int xWp, int xld, int C,
                                        bf16* __restrict__ y, int yhl, int yhr, int yWp, int yld, int sh, int sw) {
  const int vec = C >> 3;
  const int Ho = H * sh, Wo = W * sw;
  const int64_t total = (int64_t)B * Ho * Wo * vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vec);
    int64_t r = i / vec;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(b * H + ho / sh) * xWp + (wo / sw + xhl)) * xld) + cv);
    store_with_halo(y, b, Ho, Wo, yhl, yhr, yWp, yld, ho, wo, cv, v);
  }
}

// bilinear, align_corners=True: src = dst * (in-1)/(out-1)   (ATen area_pixel_compute_scale)
__global__ void upsample_bilinear_kernel(const bf16* __restrict__ x, int B, int H, int W, int xhl, int xWp, int xld,
                                         int C, bf16* __restrict__ y, int Ho, int Wo, int yhl, int yhr, int yWp,
                                         int yld, float rh, float rw, bool f16) {
  const int vec = C >> 3;
  const int64_t total = (int64_t)B * Ho * Wo * vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vec);
    int64_t r = i / vec;
    const int wo = (int)(r % Wo);
    r /= Wo;
    const int ho = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const float fh = rh * ho, fw = rw * wo;
    const int h0 = (int)fh, w0 = (int)fw;
    const int h1 = h0 + (h0 < H - 1 ? 1 : 0), w1 = w0 + (w0 < W - 1 ? 1 : 0);
    const float lh1 = fh - h0, lw1 = fw - w0;
    const float lh0 = 1.f - lh1, lw0 = 1.f - lw1;
    const bf16* base = x + (size_t)b * H * xWp * xld;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)h0 * xWp + w0 + xhl) * xld) + cv);
    const uint4 bq = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)h0 * xWp + w1 + xhl) * xld) + cv);
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)h1 * xWp + w0 + xhl) * xld) + cv);
    const uint4 d = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)h1 * xWp + w1 + xhl) * xld) + cv);
    const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {bq.x, bq.y, bq.z, bq.w}, uc[4] = {c.x, c.y, c.z, c.w},
                   ud[4] = {d.x, d.y, d.z, d.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = unpack_hr(ua[k], f16), fb = unpack_hr(ub[k], f16), fc = unpack_hr(uc[k], f16), fd = unpack_hr(ud[k], f16);
      const float r0 = lh0 * (lw0 * fa.x + lw1 * fb.x) + lh1 * (lw0 * fc.x + lw1 * fd.x);
      const float r1 = lh0 * (lw0 * fa.y + lw1 * fb.y) + lh1 * (lw0 * fc.y + lw1 * fd.y);
      o[k] = pack_hr(r0, r1, f16);
    }
    store_with_halo(y, b, Ho, Wo, yhl, yhr, yWp, yld, ho, wo, cv, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// ------------------------------------------------------------------------------------------ softmax over rows (decoder attention)
// One CTA of 256 threads per row.  Rows of up to 4096 columns (a multiple of 4) are read ONCE as float4 into registers and
// written as 8-byte packed vectors (the three-pass scalar form below read every score three times: 117 us per head of the R2DM
// attention at B = 32); the reductions keep a fixed order (thread partial -> xor-shuffle tree -> warp partials in warp order).
__global__ void __launch_bounds__(256) softmax_rows_vec_kernel(const float* __restrict__ s, bf16* __restrict__ p, int cols, bool f16) {
  const int64_t row = blockIdx.x;
  const float4* sr = reinterpret_cast<const float4*>(s + row * cols);
  uint2* pr = reinterpret_cast<uint2*>(p + row * cols);
  __shared__ float red[8];
  const int nv = cols >> 2;                       // float4 vectors in the row (<= 1024)
  float4 v[4];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < nv) {
      v[k] = __ldg(sr + i);
      mx = fmaxf(fmaxf(mx, fmaxf(v[k].x, v[k].y)), fmaxf(v[k].z, v[k].w));
    }
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < nv) {
      v[k].x = __expf(v[k].x - mx); v[k].y = __expf(v[k].y - mx); v[k].z = __expf(v[k].z - mx); v[k].w = __expf(v[k].w - mx);
      sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.f / sum;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < nv) pr[i] = make_uint2(pack_hr(v[k].x * inv, v[k].y * inv, f16), pack_hr(v[k].z * inv, v[k].w * inv, f16));
  }
}

__global__ void softmax_rows_kernel(const float* __restrict__ s, bf16* __restrict__ p, int cols, bool f16) {
  const int64_t row = blockIdx.x;
  const float* sr = s + row * cols;
  bf16* pr = p + row * cols;
  __shared__ float red[32];
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) mx = fmaxf(mx, sr[i]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) sum += __expf(sr[i] - mx);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) sum += red[i];
  const float inv = 1.f / sum;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) store_hr(pr + i, __expf(sr[i] - mx) * inv, f16);
}

// ------------------------------------------------------------------------------------------ vector quantiser
__global__ void codebook_norm_kernel(const float* __restrict__ cb, int n, int dim, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int d = 0; d < dim; ++d) s = __fadd_rn(s, __fmul_rn(cb[i * dim + d], cb[i * dim + d]));
  out[i] = s;
}

// One thread per latent pixel, codebook streamed through shared memory.  dim is fixed to 8 (embed_dim of every
// released LiDM autoencoder).  d_j = (|z|^2 + |e_j|^2) - 2 z.e_j in fp32, first minimum wins (torch.argmin).
constexpr int VQ_TILE = 1024;
__global__ void vq_kernel(const float* __restrict__ z, int C, int HW, int64_t npix, const float* __restrict__ cb,
                          const float* __restrict__ cbn, int n_embed, int quantize, const float* __restrict__ pq_w,
                          const float* __restrict__ pq_b, float scale, float* __restrict__ out, int32_t* __restrict__ idx_out) {
  __shared__ float4 sh_cb[VQ_TILE * 2];
  __shared__ float sh_n[VQ_TILE];
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = pix < npix;
  const int64_t b = active ? pix / HW : 0;
  const int64_t hw = active ? pix - b * HW : 0;
  float zv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) zv[c] = active ? __fmul_rn(scale, z[(b * C + c) * HW + hw]) : 0.f;
  float q[8];
  int best = 0;
  if (quantize) {
    float zn = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) zn = __fadd_rn(zn, __fmul_rn(zv[c], zv[c]));
    float bestd = INFINITY;
    for (int j0 = 0; j0 < n_embed; j0 += VQ_TILE) {
      __syncthreads();
      const int nt = min(VQ_TILE, n_embed - j0);
      for (int i = threadIdx.x; i < nt * 2; i += blockDim.x) sh_cb[i] = reinterpret_cast<const float4*>(cb + (size_t)j0 * 8)[i];
      for (int i = threadIdx.x; i < nt; i += blockDim.x) sh_n[i] = cbn[j0 + i];
      __syncthreads();
#pragma unroll 4
      for (int j = 0; j < nt; ++j) {
        const float4 e0 = sh_cb[2 * j], e1 = sh_cb[2 * j + 1];
        float dot = zv[0] * e0.x;
        dot = fmaf(zv[1], e0.y, dot); dot = fmaf(zv[2], e0.z, dot); dot = fmaf(zv[3], e0.w, dot);
        dot = fmaf(zv[4], e1.x, dot); dot = fmaf(zv[5], e1.y, dot); dot = fmaf(zv[6], e1.z, dot);
        dot = fmaf(zv[7], e1.w, dot);
        const float d = __fsub_rn(__fadd_rn(zn, sh_n[j]), __fmul_rn(2.f, dot));
        if (d < bestd) { bestd = d; best = j0 + j; }
      }
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float e = cb[(size_t)best * 8 + c];
      q[c] = __fadd_rn(zv[c], __fsub_rn(e, zv[c]));  // z + (z_q - z): the straight-through value the reference decodes
    }
  } else {
#pragma unroll
    for (int c = 0; c < 8; ++c) q[c] = zv[c];
  }
  if (!active) return;
  if (idx_out != nullptr) idx_out[pix] = quantize ? best : -1;
  if (pq_w == nullptr) {   // quantiser only (VectorQuantizer2.forward's z_q)
#pragma unroll
    for (int c = 0; c < 8; ++c) out[(b * C + c) * HW + hw] = q[c];
    return;
  }
  // post_quant_conv: 1x1 conv embed_dim -> z_channels (fp32)
  for (int co = 0; co < C; ++co) {
    float acc = __ldg(pq_b + co);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc = fmaf(__ldg(pq_w + co * 8 + c), q[c], acc);
    out[(b * C + co) * HW + hw] = acc;
  }
}

// ------------------------------------------------------------------------------------------ timestep embedding MLP
// tmp[r][j] = SiLU(b0[j] + sum_i w0[j][i] * temb(t_r)[i]),  temb = [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(1e4) i / half)
__global__ void time_embed_l0_kernel(const int64_t* __restrict__ t, int t_stride, int model_ch, const float* __restrict__ w0,
                                     const float* __restrict__ b0, int ted, float* __restrict__ tmp, int style) {
  extern __shared__ float te[];
  const int r = blockIdx.y;
  const int half = model_ch / 2;
  const float tv = (float)t[(size_t)r * t_stride];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    if (style == 0) {
      const float f = expf(-logf(10000.f) * (float)i / (float)half);
      const float a = tv * f;
      te[i] = cosf(a);
      te[half + i] = sinf(a);
    } else {      // SinusoidalPositionalEmbedding (unets/ops.py:14-27): exp(h * i), h = -ln(P) / (half - 1), float64 h as numpy
      const float f = expf((float)(-9.210340371976184 / (double)(half - 1)) * (float)i);
      const float a = tv * f;
      te[i] = sinf(a);
      te[half + i] = cosf(a);
    }
  }
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ted) return;
  float acc = b0[j];
  const float* wr = w0 + (size_t)j * model_ch;
  for (int i = 0; i < model_ch; ++i) acc = fmaf(wr[i], te[i], acc);
  tmp[(size_t)r * ted + j] = acc / (1.f + expf(-acc));
}

// out[r][n] = act(b[n] + sum_k w[n][k] x[r][k]); one warp per output n, loops over rows r.
__global__ void linear_rows_kernel(const float* __restrict__ x, int nt, int K, const float* __restrict__ w,
                                   const float* __restrict__ bvec, int N, float* __restrict__ out, int silu,
                                   const float* __restrict__ rowbias) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* wr = w + (size_t)n * K;
  for (int r = 0; r < nt; ++r) {
    const float* xr = x + (size_t)r * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(__ldg(wr + k), xr[k], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      acc += bvec[n];
      if (rowbias != nullptr) acc += rowbias[(size_t)r * N + n];
      out[(size_t)r * N + n] = silu ? acc / (1.f + expf(-acc)) : acc;
    }
  }
}

// ------------------------------------------------------------------------------------------ weight packing
// w: fp32 [cout][cin][kh][kw] -> bf16 [n_alloc][k_alloc], out[row'][tap*cin + c] with row' = row_perm ? perm[row] : row,
// rows < n_scaled_rows (after permutation) multiplied by row_scale.  Padding rows/cols are zero.
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int cout, int cin, int kh, int kw, int n_alloc,
                                        int k_alloc, const int* __restrict__ row_perm, float row_scale,
                                        int n_scaled_rows, bf16* __restrict__ out, bool f16) {
  const int64_t total = (int64_t)n_alloc * k_alloc;
  const int taps = kh * kw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % k_alloc);
    const int rowp = (int)(i / k_alloc);
    float v = 0.f;
    if (rowp < cout && k < taps * cin) {
      const int row = row_perm != nullptr ? row_perm[rowp] : rowp;   // source row feeding packed row rowp
      const int tap = k / cin, c = k - tap * cin;
      v = w[((size_t)row * cin + c) * taps + tap];
      if (rowp < n_scaled_rows) v *= row_scale;
    }
    store_hr(out + i, v, f16);
  }
}

__global__ void scale_rows_kernel(float* __restrict__ v, int n, float s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] *= s;
}

// dec (B,2,HW) -> out (B,1,HW): out = dec[:,1] < 0 ? -1 : dec[:,0]   (autoencoder.py:298-301)
__global__ void mask_select_kernel(const float* __restrict__ dec, int B, int HW, float* __restrict__ out) {
  const int64_t total = (int64_t)B * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, hw = i - b * HW;
    const float v = dec[(b * 2) * HW + hw], m = dec[(b * 2 + 1) * HW + hw];
    out[i] = m < 0.f ? -1.f : v;
  }
}

__global__ void f32_nchw_to_nhwc_bf16_kernel(const float* __restrict__ x, int B, int C, int H, int W, bf16* __restrict__ y,
                                             int hl, int hr, int Wp, int ld) {
  const int64_t total = (int64_t)B * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int w = (int)(r % W);
    r /= W;
    const int h = (int)(r % H);
    const int b = (int)(r / H);
    const bf16 v = __float2bfloat16(x[(((int64_t)b * C + c) * H + h) * W + w]);
    const size_t rowbase = (size_t)(b * H + h) * Wp;
    y[(rowbase + w + hl) * ld + c] = v;
    if (w < hr) y[(rowbase + W + hl + w) * ld + c] = v;
    if (w >= W - hl) y[(rowbase + (w - (W - hl))) * ld + c] = v;
  }
}

__global__ void nhwc_bf16_to_f32_nchw_kernel(const bf16* __restrict__ x, int B, int C, int H, int W, int hl, int Wp,
                                             int ld, float* __restrict__ y) {
  const int64_t total = (int64_t)B * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    int64_t r = i / W;
    const int h = (int)(r % H);
    r /= H;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    y[i] = __bfloat162float(x[((size_t)(b * H + h) * Wp + (w + hl)) * ld + c]);
  }
}

__global__ void copy_with_halo_kernel(const bf16* __restrict__ x, int B, int H, int W, int xhl, int xWp, int xld, int C,
                                      bf16* __restrict__ y, int yhl, int yhr, int yWp, int yld) {
  const int vec = C >> 3;
  const int64_t total = (int64_t)B * H * W * vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % vec);
    int64_t r = i / vec;
    const int w = (int)(r % W);
    r /= W;
    const int h = (int)(r % H);
    const int b = (int)(r / H);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + ((size_t)(b * H + h) * xWp + (w + xhl)) * xld) + cv);
    store_with_halo(y, b, H, W, yhl, yhr, yWp, yld, h, w, cv, v);
  }
}

inline int grid_for(int64_t total, int threads) {
  int64_t g = (total + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

void launch_ddim_step(const float* x, const float* eps, const float* noise, const float* coef_dev, float* x_prev,
                      float* pred_x0, int64_t n, cudaStream_t s) {
  LIDM_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)eps & 15) == 0 && ((uintptr_t)x_prev & 15) == 0 &&
                   ((uintptr_t)pred_x0 & 15) == 0 && ((uintptr_t)noise & 15) == 0,
               "ddim_step: pointers must be 16-byte aligned");
  ddim_step_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, s>>>(x, eps, noise, coef_dev, x_prev, pred_x0, n);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_ddpm_step(const float* x, const float* eps, const float* noise, const float* coef, int B, int64_t n_per_sample, int clip,
                      float* x_prev, float* x_recon, cudaStream_t s) {
  dim3 grid((unsigned)std::min<int64_t>((n_per_sample + 255) / 256, 1024), B);
  ddpm_step_kernel<<<grid, 256, 0, s>>>(x, eps, noise, coef, n_per_sample, clip, x_prev, x_recon);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_backproject(const float* img, int B, int H, int W, float fov_up_deg, float fov_down_deg, float dmin,
                        float dmax, float depth_scale, int log_scale, int input_is_unit, float* xyz, uint8_t* mask,
                        cudaStream_t s) {
  LIDM_REQUIRE(B > 0 && H > 0 && W > 0, "backproject: empty image");
  LIDM_REQUIRE(W % 4 == 0 && ((uintptr_t)img & 15) == 0 && ((uintptr_t)xyz & 15) == 0 && ((uintptr_t)mask & 3) == 0,
               "backproject: W must be a multiple of 4 and buffers 16-byte aligned");
  const double kPi = 3.14159265358979323846;
  const double fu = (double)fov_up_deg / 180.0 * kPi, fd = (double)fov_down_deg / 180.0 * kPi;
  const int threads = 128;
  dim3 grid(cdiv(W, threads * 4), H, B);
  backproject_kernel<<<grid, threads, threads * 4 * 2 * sizeof(float), s>>>(img, H, W, fu, fd, dmin, dmax, depth_scale,
                                                                             log_scale, input_is_unit, xyz, mask);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_im2col_nchw_f32(const float* x, int B, int C, int H, int W, int kh, int kw, int pl, int pt, bf16* out,
                            int kpad, cudaStream_t s, bool f16, bool zero_w) {
  LIDM_REQUIRE(kh * kw * C <= kpad && kpad % 8 == 0 && pl < W && kw - 1 - pl < W, "im2col: kpad too small / not a multiple of 8");
  const int64_t total = (int64_t)B * H * W * (kpad / 8);
  im2col_nchw_f32_kernel<<<grid_for(total, 256), 256, 0, s>>>(x, B, C, H, W, kh, kw, pl, pt, out, kpad, f16, zero_w);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_im2col_nhwc(const View& x, int kh, int kw, int stride, int pl, int pt, int Ho, int Wo, bf16* out,
                        cudaStream_t s, int stride_w) {
  LIDM_REQUIRE(x.C % 8 == 0, "im2col: C % 8");
  const int64_t total = (int64_t)x.B * Ho * Wo * kh * kw * (x.C / 8);
  im2col_nhwc_kernel<<<grid_for(total, 256), 256, 0, s>>>(x.p, x.B, x.H, x.W, x.hl, x.Wp(), x.ld, x.C, kh, kw, stride,
                                                          stride_w > 0 ? stride_w : stride, pl, pt, Ho, Wo, out);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_upsample_nearest2x(const View& x, const View& y, cudaStream_t s) {
  LIDM_REQUIRE(y.H == 2 * x.H && y.W == 2 * x.W && y.C == x.C && y.B == x.B, "nearest upsample shapes");
  const int64_t total = (int64_t)y.B * y.H * y.W * (y.C / 8);
  upsample_nearest_kernel<<<grid_for(total, 256), 256, 0, s>>>(x.p, x.B, x.H, x.W, x.hl, x.Wp(), x.ld, x.C, y.p, y.hl,
                                                               y.hr, y.Wp(), y.ld, 2, 2);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_upsample_bilinear(const View& x, const View& y, cudaStream_t s) {
  LIDM_REQUIRE(y.C == x.C && y.B == x.B && y.H >= x.H && y.W >= x.W && x.f16 == y.f16, "bilinear upsample shapes");
  const float rh = y.H > 1 ? (float)(x.H - 1) / (float)(y.H - 1) : 0.f;
  const float rw = y.W > 1 ? (float)(x.W - 1) / (float)(y.W - 1) : 0.f;
  const int64_t total = (int64_t)y.B * y.H * y.W * (y.C / 8);
  upsample_bilinear_kernel<<<grid_for(total, 256), 256, 0, s>>>(x.p, x.B, x.H, x.W, x.hl, x.Wp(), x.ld, x.C, y.p, y.H,
                                                                y.W, y.hl, y.hr, y.Wp(), y.ld, rh, rw, x.f16);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_copy_with_halo(const View& x, const View& y, cudaStream_t s) {
  LIDM_REQUIRE(y.C == x.C && y.B == x.B && y.H == x.H && y.W == x.W, "copy shapes");
  const int64_t total = (int64_t)x.B * x.H * x.W * (x.C / 8);
  copy_with_halo_kernel<<<grid_for(total, 256), 256, 0, s>>>(x.p, x.B, x.H, x.W, x.hl, x.Wp(), x.ld, x.C, y.p, y.hl,
                                                             y.hr, y.Wp(), y.ld);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_softmax_rows(const float* sc, bf16* p, int64_t rows, int cols, cudaStream_t st, bool f16) {
  if (cols % 4 == 0 && cols <= 4096 && (reinterpret_cast<uintptr_t>(sc) & 15) == 0 && (reinterpret_cast<uintptr_t>(p) & 7) == 0)
    softmax_rows_vec_kernel<<<(unsigned)rows, 256, 0, st>>>(sc, p, cols, f16);
  else
    softmax_rows_kernel<<<(unsigned)rows, 256, 0, st>>>(sc, p, cols, f16);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_codebook_norm(const float* codebook, int n_embed, int dim, float* out, cudaStream_t s) {
  codebook_norm_kernel<<<cdiv(n_embed, 256), 256, 0, s>>>(codebook, n_embed, dim, out);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_vq(const float* z, int B, int C, int HW, const float* codebook, const float* cb_norm, int n_embed,
               int quantize, const float* pq_w, const float* pq_b, float scale, float* out, int32_t* idx,
               cudaStream_t s) {
  LIDM_REQUIRE(C == 8, "vq: embed_dim/z_channels must be 8");
  const int64_t npix = (int64_t)B * HW;
  vq_kernel<<<cdiv(npix, 128), 128, 0, s>>>(z, C, HW, npix, codebook, cb_norm, n_embed, quantize, pq_w, pq_b, scale, out,
                                            idx);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_time_embed(const int64_t* t_dev, int nt, int model_ch, const float* w0, const float* b0, const float* w2,
                       const float* b2, int ted, float* tmp, float* emb_silu, cudaStream_t s, int t_stride,
                       const float* rowbias, int style) {
  dim3 g0(cdiv(ted, 128), nt);
  time_embed_l0_kernel<<<g0, 128, model_ch * sizeof(float), s>>>(t_dev, t_stride, model_ch, w0, b0, ted, tmp, style);
  LIDM_CUDA_CHECK(cudaGetLastError());
  // emb = Linear(tmp) (+ the layout encoder's xf_proj row); every consumer applies SiLU first (openaimodel.py:222-223)
  // so store SiLU(emb)
  linear_rows_kernel<<<cdiv(ted, 8), 256, 0, s>>>(tmp, nt, ted, w2, b2, ted, emb_silu, 1, rowbias);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_linear_rows(const float* x, int nt, int K, const float* w, const float* b, int N, float* out,
                        cudaStream_t s) {
  linear_rows_kernel<<<cdiv(N, 8), 256, 0, s>>>(x, nt, K, w, b, N, out, 0, nullptr);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_pack_conv_weight(const float* w, int cout, int cin, int kh, int kw, int n_alloc, int k_alloc,
                             const int* row_perm, const float* /*unused*/, float row_scale, int n_scaled_rows,
                             bf16* out, cudaStream_t s, bool f16) {
  const int64_t total = (int64_t)n_alloc * k_alloc;
  pack_conv_weight_kernel<<<grid_for(total, 256), 256, 0, s>>>(w, cout, cin, kh, kw, n_alloc, k_alloc, row_perm,
                                                               row_scale, n_scaled_rows, out, f16);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_mask_select(const float* dec, int B, int HW, float* out, cudaStream_t s) {
  mask_select_kernel<<<grid_for((int64_t)B * HW, 256), 256, 0, s>>>(dec, B, HW, out);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_f32_to_nhwc_bf16(const float* x, int B, int C, int HW, const View& y, cudaStream_t s) {
  LIDM_REQUIRE(y.H * y.W == HW && y.C == C && y.B == B, "layout conversion shapes");
  f32_nchw_to_nhwc_bf16_kernel<<<grid_for((int64_t)B * HW * C, 256), 256, 0, s>>>(x, B, C, y.H, y.W, y.p, y.hl, y.hr,
                                                                                  y.Wp(), y.ld);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_nhwc_bf16_to_f32_nchw(const View& x, float* y, cudaStream_t s) {
  nhwc_bf16_to_f32_nchw_kernel<<<grid_for((int64_t)x.B * x.H * x.W * x.C, 256), 256, 0, s>>>(x.p, x.B, x.C, x.H, x.W,
                                                                                             x.hl, x.Wp(), x.ld, y);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

}  // namespace lidm
