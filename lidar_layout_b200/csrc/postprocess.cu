// Sample post-processing on the device (the step right after the sampling path in the reference's script,
// scripts/sample.py:29-45,129-162):
//   * custom_to_pil: uint8 image of (clip(x,-1,1)+1)/2*255, truncated like ndarray.astype(np.uint8)
//   * custom_to_pcd / range2pcd's `pcd[mask, :]`: stream compaction of the valid points of every range image, in
//     row-major pixel order (the order numpy boolean indexing produces), one (N_b, 3) fp32 block per sample.
// HBM-bound, one read of the inputs and one write of the outputs.
#include "common.h"

namespace lidm {

namespace {

__global__ void to_uint8_image_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    float v[4];
    if (i + 4 <= n) {
      const float4 t = *reinterpret_cast<const float4*>(x + i);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
      for (int k = 0; k < 4; ++k) v[k] = (i + k < n) ? x[i + k] : 0.f;
    }
    uint8_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float r = fminf(fmaxf(v[k], -1.f), 1.f);
      r = __fdiv_rn(__fadd_rn(r, 1.f), 2.f);
      r = __fmul_rn(255.f, r);
      o[k] = (uint8_t)(int)r;          // r in [0, 255]: truncation, as astype(np.uint8)
    }
    if (i + 4 <= n) *reinterpret_cast<uchar4*>(y + i) = make_uchar4(o[0], o[1], o[2], o[3]);
    else for (int k = 0; k < 4 && i + k < n; ++k) y[i + k] = o[k];
  }
}

// One CTA per sample walks its HW pixels in chunks of blockDim.x * 4; a block-wide exclusive scan of the per-thread
// valid counts gives every point its slot, so the output order is the pixel order.
__global__ void __launch_bounds__(1024)
compact_points_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ mask, int HW,
                      float* __restrict__ points, int32_t* __restrict__ counts) {
  __shared__ int warp_sums[32];
  __shared__ int base_s;
  const int b = blockIdx.x;
  const float* x = xyz + (size_t)b * 3 * HW;
  const uint8_t* m = mask + (size_t)b * HW;
  float* out = points + (size_t)b * HW * 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int p0 = 0; p0 < HW; p0 += blockDim.x * 4) {
    const int p = p0 + threadIdx.x * 4;
    uint8_t mk[4] = {0, 0, 0, 0};
    if (p + 4 <= HW) {
      const uchar4 t = *reinterpret_cast<const uchar4*>(m + p);
      mk[0] = t.x; mk[1] = t.y; mk[2] = t.z; mk[3] = t.w;
    } else {
      for (int k = 0; k < 4; ++k) mk[k] = (p + k < HW) ? m[p + k] : 0;
    }
    const int cnt = (mk[0] != 0) + (mk[1] != 0) + (mk[2] != 0) + (mk[3] != 0);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int v = lane < nwarps ? warp_sums[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      warp_sums[lane] = v;   // inclusive prefix over warps
    }
    __syncthreads();
    const int base = base_s;
    int slot = base + (warp > 0 ? warp_sums[warp - 1] : 0) + incl - cnt;
    const int chunk_total = warp_sums[nwarps - 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (mk[k]) {
        out[(size_t)slot * 3 + 0] = x[p + k];
        out[(size_t)slot * 3 + 1] = x[HW + p + k];
        out[(size_t)slot * 3 + 2] = x[2 * (size_t)HW + p + k];
        ++slot;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) base_s = base + chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b] = base_s;
}

}  // namespace

void launch_to_uint8_image(const float* x, uint8_t* y, int64_t n, cudaStream_t s) {
  LIDM_REQUIRE(n > 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 3) == 0, "to_uint8_image: alignment");
  int64_t g = (n / 4 + 255) / 256;
  if (g < 1) g = 1;
  if (g > 148 * 16) g = 148 * 16;
  to_uint8_image_kernel<<<(int)g, 256, 0, s>>>(x, y, n);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_compact_points(const float* xyz, const uint8_t* mask, int B, int HW, float* points, int32_t* counts,
                           cudaStream_t s) {
  LIDM_REQUIRE(B > 0 && HW > 0 && ((uintptr_t)mask & 3) == 0 && HW % 4 == 0, "compact_points: HW must be a multiple of 4");
  compact_points_kernel<<<B, 1024, 0, s>>>(xyz, mask, HW, points, counts);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

}  // namespace lidm
