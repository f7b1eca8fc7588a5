// Nearest-neighbour squared distances between two point sets (the forward of the reference's Chamfer-distance extension,
// lidm/eval/modules/chamfer3D/chamfer3D.cu:12-155 and chamfer2D/chamfer2D.cu:12-145, called from
// lidm/eval/metric_utils.py:414-440): for every point of set A the squared distance to, and the index of, its nearest
// point in set B.  Brute force like the reference, laid out for the B200: one thread owns two query points in registers,
// a CTA walks set B in 1024-point tiles staged in shared memory as x / y / z planes (every lane reads the same address:
// a broadcast), grid = (query chunks, batch).  The distance is evaluated as (dx*dx + dy*dy) + dz*dz with every
// operation rounded on its own (no FMA contraction), so the result is the same fp32 number a numpy restatement
// produces; ties go to the lowest index (strict <, ascending scan), as in the reference.
#include "common.h"

namespace lidm {

namespace {

constexpr int TILE = 1024;
constexpr int THREADS = 256;
constexpr int QPT = 2;   // query points per thread

template <int DIM>
__global__ void __launch_bounds__(THREADS)
nn_dist_kernel(const float* __restrict__ a, int n, const float* __restrict__ b, int m, float* __restrict__ dist,
               int32_t* __restrict__ idx) {
  __shared__ float sx[TILE], sy[TILE], sz[DIM == 3 ? TILE : 1];
  const int bi = blockIdx.y;
  const float* ap = a + (size_t)bi * n * DIM;
  const float* bp = b + (size_t)bi * m * DIM;
  float qx[QPT], qy[QPT], qz[QPT], best[QPT];
  int besti[QPT];
  int q[QPT];
#pragma unroll
  for (int t = 0; t < QPT; ++t) {
    q[t] = (blockIdx.x * QPT + t) * THREADS + threadIdx.x;
    const int qq = q[t] < n ? q[t] : n - 1;
    qx[t] = __ldg(ap + (size_t)qq * DIM);
    qy[t] = __ldg(ap + (size_t)qq * DIM + 1);
    qz[t] = DIM == 3 ? __ldg(ap + (size_t)qq * DIM + 2) : 0.f;
    best[t] = 0.f;
    besti[t] = 0;
  }
  for (int k0 = 0; k0 < m; k0 += TILE) {
    const int cnt = min(TILE, m - k0);
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += THREADS) {
      const float* p = bp + (size_t)(k0 + j) * DIM;
      sx[j] = __ldg(p);
      sy[j] = __ldg(p + 1);
      if (DIM == 3) sz[j] = __ldg(p + 2);
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      const float bx = sx[k], by = sy[k], bz = DIM == 3 ? sz[k] : 0.f;
#pragma unroll
      for (int t = 0; t < QPT; ++t) {
        const float dx = __fsub_rn(bx, qx[t]), dy = __fsub_rn(by, qy[t]);
        float d = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (DIM == 3) {
          const float dz = __fsub_rn(bz, qz[t]);
          d = __fadd_rn(d, __fmul_rn(dz, dz));
        }
        if ((k0 + k == 0) || d < best[t]) {
          best[t] = d;
          besti[t] = k0 + k;
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < QPT; ++t) {
    if (q[t] < n) {
      dist[(size_t)bi * n + q[t]] = best[t];
      idx[(size_t)bi * n + q[t]] = besti[t];
    }
  }
}

}  // namespace

void launch_nn_dist(const float* a, int n, const float* b, int m, int B, int dim, float* dist, int32_t* idx, cudaStream_t s) {
  LIDM_REQUIRE(a && b && dist && idx && B > 0 && n > 0 && m > 0 && (dim == 2 || dim == 3), "nearest-neighbour distance arguments");
  dim3 grid((n + THREADS * QPT - 1) / (THREADS * QPT), B);
  if (dim == 3) nn_dist_kernel<3><<<grid, THREADS, 0, s>>>(a, n, b, m, dist, idx);
  else nn_dist_kernel<2><<<grid, THREADS, 0, s>>>(a, n, b, m, dist, idx);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

}  // namespace lidm
