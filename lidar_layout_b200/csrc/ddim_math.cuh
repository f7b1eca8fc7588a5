// DDIM x_{t-1} update, bit-faithful to the reference's fp32 op order
// (lidm/models/diffusion/ddim.py:196-205): every intermediate is rounded like the separate torch ops
// (no FMA contraction), so given the same eps the result equals the reference's to the last bit.
#pragma once
#include <cuda_runtime.h>

namespace lidm {

// coef = {a_t, a_prev, sigma_t, sqrt_one_minus_at, temperature}
__device__ __forceinline__ void ddim_update(float x, float e_t, float noise, const float (&coef)[5], float& x_prev,
                                            float& pred_x0) {
  const float a_t = coef[0], a_prev = coef[1], sigma_t = coef[2], sqrt_one_minus_at = coef[3], temp = coef[4];
  // pred_x0 = (x - sqrt_one_minus_at * e_t) / a_t.sqrt()
  pred_x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(sqrt_one_minus_at, e_t)), __fsqrt_rn(a_t));
  // dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
  const float dir_xt =
      __fmul_rn(__fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, a_prev), __fmul_rn(sigma_t, sigma_t))), e_t);
  // noise = sigma_t * noise_like(...) * temperature
  const float nz = __fmul_rn(__fmul_rn(sigma_t, noise), temp);
  // x_prev = a_prev.sqrt() * pred_x0 + dir_xt + noise
  x_prev = __fadd_rn(__fadd_rn(__fmul_rn(__fsqrt_rn(a_prev), pred_x0), dir_xt), nz);
}

}  // namespace lidm
