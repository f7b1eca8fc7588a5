// Nearest-neighbour squared distances between two point sets (the forward of the reference's Chamfer-distance extension,
// lidm/eval/modules/chamfer3D/chamfer3D.cu:12-155 and chamfer2D/chamfer2D.cu:12-145, called from
// lidm/eval/metric_utils.py:414-440): for every point of set A the squared distance to, and the index of, its nearest
// point in set B.  Brute force like the reference, laid out for the B200: one thread owns two query points in registers,
// a CTA walks set B in 1024-point tiles staged in shared memory as x / y / z planes (every lane reads the same address:
// a broadcast), grid = (query chunks, batch).  The distance is evaluated as (dx*dx + dy*dy) + dz*dz with every
// operation rounded on its own (no FMA contraction), so the result is the same fp32 number a numpy restatement
// produces; ties go to the lowest index (strict <, ascending scan), as in the reference.
#include <algorithm>

#include "common.h"

namespace lidm {

namespace {

constexpr int TILE = 1024;
constexpr int THREADS = 256;
constexpr int QPT = 2;   // query points per thread

// FMA: the distance contracted the way nvcc compiles the reference's `dx*dx + dy*dy + dz*dz` by default
// (FMUL dy*dy, FFMA dx*dx + ., FFMA dz*dz + . in the SASS of the reference extension: fma(dz, dz, fma(dx, dx, dy*dy))): bit-identical to the reference extension built for sm_100 (tests/test_gpu_eval_ref.py);
// !FMA: every operation rounded on its own (what a numpy restatement computes).
template <int DIM, bool FMA>
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
        float d = FMA ? __fmaf_rn(dx, dx, __fmul_rn(dy, dy)) : __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (DIM == 3) {
          const float dz = __fsub_rn(bz, qz[t]);
          d = FMA ? __fmaf_rn(dz, dz, d) : __fadd_rn(d, __fmul_rn(dz, dz));
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

// --------------------------------------------------------------------------------------------------- Chamfer backward
// NmDistanceGradKernel (chamfer3D.cu:155-171 / chamfer2D.cu): d dist_a[j] / d a[j] = 2 (a[j] - b[idx[j]]), scattered to both
// sets; the scatter into the other set uses atomics like the reference (several points may share a nearest neighbour).
template <int DIM>
__global__ void chamfer_grad_kernel(const float* __restrict__ a, int n, const float* __restrict__ b, int m,
                                    const float* __restrict__ grad_dist, const int32_t* __restrict__ idx, float* __restrict__ grad_a,
                                    float* __restrict__ grad_b) {
  const int bi = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int j2 = idx[(size_t)bi * n + j];
    const float g = grad_dist[(size_t)bi * n + j] * 2;
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      const float v = g * (a[((size_t)bi * n + j) * DIM + c] - b[((size_t)bi * m + j2) * DIM + c]);
      atomicAdd(grad_a + ((size_t)bi * n + j) * DIM + c, v);
      atomicAdd(grad_b + ((size_t)bi * m + j2) * DIM + c, -v);
    }
  }
}

// --------------------------------------------------------------------------------------------------- EMD (auction)
// The auction algorithm of the reference's EMD extension (lidm/eval/modules/emd/emd_cuda.cu:23-284, author Minghua Liu):
// every iteration lists the unassigned points, each bids for the object maximising 3 - |p - q| - price with increment
// best - second best + eps, every object takes its highest bidder (evicting the previous owner) and raises its price; the
// last iteration assigns every remaining point to its bid.  Same arithmetic (the value is formed in double from the float
// square root, as `3.0 - sqrtf(..) - price` does), different mapping: unassigned points are compacted in ascending order by
// a block scan, ONE WARP bids for one point (lanes stride over the objects staged in shared memory), ties go to the lowest
// object index exactly as the reference's ascending strict-> scan, and equal top bidders are resolved deterministically
// (highest point index) where the reference lets the last writer win.
constexpr int EMD_TILE = 1024;

__global__ void __launch_bounds__(1024) emd_unassigned_kernel(const int32_t* __restrict__ assignment, int n, int32_t* __restrict__ unass_idx,
                                                              int32_t* __restrict__ unass_cnt, int32_t* __restrict__ max_idx) {
  __shared__ int warp_sum[32];
  __shared__ int base;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int j0 = 0; j0 < n; j0 += 1024) {
    const int j = j0 + threadIdx.x;
    if (j < n) max_idx[(size_t)b * n + j] = -1;
    const int flag = (j < n && assignment[(size_t)b * n + j] == -1) ? 1 : 0;
    int incl = flag;
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int v = warp_sum[lane];
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
      warp_sum[lane] = v;                       // inclusive sums of the warp totals
    }
    __syncthreads();
    const int before = base + (warp ? warp_sum[warp - 1] : 0) + incl - flag;
    if (flag) unass_idx[(size_t)b * n + before] = j;
    __syncthreads();
    if (threadIdx.x == 0) base += warp_sum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) unass_cnt[b] = base;
}

__device__ __forceinline__ float atomic_max_float(float* address, float val) {
  int ret = __float_as_int(*address);
  while (val > __int_as_float(ret)) {
    const int old = ret;
    if ((ret = atomicCAS(reinterpret_cast<int*>(address), old, __float_as_int(val))) == old) break;
  }
  return __int_as_float(ret);
}

template <bool FMA>
__global__ void __launch_bounds__(256) emd_bid_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                      const float* __restrict__ price, int n, float eps,
                                                      const int32_t* __restrict__ unass_idx, const int32_t* __restrict__ unass_cnt,
                                                      int32_t* __restrict__ bid, float* __restrict__ bid_inc, float* __restrict__ max_inc) {
  __shared__ float sx[EMD_TILE], sy[EMD_TILE], sz[EMD_TILE], sp[EMD_TILE];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cnt = unass_cnt[b];
  if ((int)blockIdx.x * 8 >= cnt) return;                    // block-uniform
  const int u = blockIdx.x * 8 + warp;
  const bool active = u < cnt;
  const int pid = active ? unass_idx[(size_t)b * n + u] : 0;
  const float x1 = xyz1[((size_t)b * n + pid) * 3], y1 = xyz1[((size_t)b * n + pid) * 3 + 1], z1 = xyz1[((size_t)b * n + pid) * 3 + 2];
  float best = -1e9f, better = -1e9f;
  int best_i = -1;
  for (int k0 = 0; k0 < n; k0 += EMD_TILE) {
    const int tile = min(EMD_TILE, n - k0);
    __syncthreads();
    for (int j = threadIdx.x; j < tile; j += 256) {
      const float* q = xyz2 + ((size_t)b * n + k0 + j) * 3;
      sx[j] = q[0]; sy[j] = q[1]; sz[j] = q[2];
      sp[j] = price[(size_t)b * n + k0 + j];
    }
    __syncthreads();
    if (active) {
      for (int k = lane; k < tile; k += 32) {
        const float dx = sx[k] - x1, dy = sy[k] - y1, dz = sz[k] - z1;
        const float s2 = FMA ? __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)))
                             : __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        const float d = (float)(3.0 - (double)sqrtf(s2) - (double)sp[k]);
        if (d > best) { better = best; best = d; best_i = k0 + k; }
        else if (d > better) better = d;
      }
    }
  }
  if (!active) return;
  // merge the lanes' (best, second best, index): the better value wins, equal values go to the lower object index
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, better, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    const bool take = oi >= 0 && (best_i < 0 || ob > best || (ob == best && oi < best_i));
    if (take) { better = fmaxf(best, os); best = ob; best_i = oi; }
    else better = fmaxf(better, ob);
  }
  if (lane == 0) {
    const float inc = best - better + eps;
    bid[(size_t)b * n + pid] = best_i;
    bid_inc[(size_t)b * n + pid] = inc;
    atomic_max_float(max_inc + (size_t)b * n + best_i, inc);
  }
}

__global__ void emd_getmax_kernel(int n, const int32_t* __restrict__ assignment, const int32_t* __restrict__ bid,
                                  const float* __restrict__ bid_inc, const float* __restrict__ max_inc, int32_t* __restrict__ max_idx) {
  const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || assignment[(size_t)b * n + j] != -1) return;
  const int o = bid[(size_t)b * n + j];
  const float bi = bid_inc[(size_t)b * n + j], mi = max_inc[(size_t)b * n + o];
  if (bi - 1e-6 <= mi && mi <= bi + 1e-6) atomicMax(max_idx + (size_t)b * n + o, j);
}

__global__ void emd_assign_kernel(int n, int32_t* __restrict__ assignment, int32_t* __restrict__ assignment_inv, float* __restrict__ price,
                                  const int32_t* __restrict__ bid, const float* __restrict__ bid_inc, float* __restrict__ max_inc,
                                  const int32_t* __restrict__ max_idx, int last) {
  const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || assignment[(size_t)b * n + j] != -1) return;
  const int o = bid[(size_t)b * n + j];
  if (last || max_idx[(size_t)b * n + o] == j) {
    const int prev = assignment_inv[(size_t)b * n + o];
    if (!last && prev != -1) assignment[(size_t)b * n + prev] = -1;
    assignment_inv[(size_t)b * n + o] = j;
    assignment[(size_t)b * n + j] = o;
    if (last) atomicAdd(price + (size_t)b * n + o, bid_inc[(size_t)b * n + j]);
    else price[(size_t)b * n + o] += bid_inc[(size_t)b * n + j];
    max_inc[(size_t)b * n + o] = -1e9f;
  }
}

__global__ void emd_dist_kernel(int n, const float* __restrict__ xyz1, const float* __restrict__ xyz2, const int32_t* __restrict__ assignment,
                                float* __restrict__ dist) {
  const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int k = assignment[(size_t)b * n + j];
  const float dx = xyz1[((size_t)b * n + j) * 3] - xyz2[((size_t)b * n + k) * 3];
  const float dy = xyz1[((size_t)b * n + j) * 3 + 1] - xyz2[((size_t)b * n + k) * 3 + 1];
  const float dz = xyz1[((size_t)b * n + j) * 3 + 2] - xyz2[((size_t)b * n + k) * 3 + 2];
  dist[(size_t)b * n + j] = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__global__ void emd_grad_kernel(int n, const float* __restrict__ xyz1, const float* __restrict__ xyz2, const float* __restrict__ grad_dist,
                                const int32_t* __restrict__ assignment, float* __restrict__ grad_xyz) {
  const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int k = assignment[(size_t)b * n + j];
  const float g = grad_dist[(size_t)b * n + j] * 2;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    grad_xyz[((size_t)b * n + j) * 3 + c] += g * (xyz1[((size_t)b * n + j) * 3 + c] - xyz2[((size_t)b * n + k) * 3 + c]);
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

void launch_chamfer_grad(const float* a, int n, const float* b, int m, int B, int dim, const float* grad_dist, const int32_t* idx,
                         float* grad_a, float* grad_b, cudaStream_t s) {
  dim3 grid(std::min((n + 255) / 256, 64), B);
  if (dim == 3) chamfer_grad_kernel<3><<<grid, 256, 0, s>>>(a, n, b, m, grad_dist, idx, grad_a, grad_b);
  else chamfer_grad_kernel<2><<<grid, 256, 0, s>>>(a, n, b, m, grad_dist, idx, grad_a, grad_b);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_emd_forward(const float* xyz1, const float* xyz2, int B, int n, float eps, int iters, float* dist, int32_t* assignment,
                        void* workspace, cudaStream_t s, bool fma) {
  // workspace: assignment_inv, bid, unass_idx, max_idx (int32 B*n each), price, bid_inc, max_inc (float B*n each), unass_cnt (B)
  const size_t N = (size_t)B * n;
  int32_t* assignment_inv = reinterpret_cast<int32_t*>(workspace);
  int32_t* bid = assignment_inv + N;
  int32_t* unass_idx = bid + N;
  int32_t* max_idx = unass_idx + N;
  float* price = reinterpret_cast<float*>(max_idx + N);
  float* bid_inc = price + N;
  float* max_inc = bid_inc + N;
  int32_t* unass_cnt = reinterpret_cast<int32_t*>(max_inc + N);
  fill_i32_kernel<<<296, 256, 0, s>>>(assignment, (int64_t)N, -1);
  fill_i32_kernel<<<296, 256, 0, s>>>(assignment_inv, (int64_t)N, -1);
  LIDM_CUDA_CHECK(cudaMemsetAsync(bid, 0, 3 * N * sizeof(int32_t), s));          // bid, unass_idx, max_idx
  LIDM_CUDA_CHECK(cudaMemsetAsync(price, 0, 3 * N * sizeof(float), s));          // price, bid_inc, max_inc (zeros, as the reference)
  dim3 gpts((n + 255) / 256, B), gbid((n + 7) / 8, B);
  for (int it = 0; it < iters; ++it) {
    emd_unassigned_kernel<<<B, 1024, 0, s>>>(assignment, n, unass_idx, unass_cnt, max_idx);
    if (fma) emd_bid_kernel<true><<<gbid, 256, 0, s>>>(xyz1, xyz2, price, n, eps, unass_idx, unass_cnt, bid, bid_inc, max_inc);
    else emd_bid_kernel<false><<<gbid, 256, 0, s>>>(xyz1, xyz2, price, n, eps, unass_idx, unass_cnt, bid, bid_inc, max_inc);
    emd_getmax_kernel<<<gpts, 256, 0, s>>>(n, assignment, bid, bid_inc, max_inc, max_idx);
    emd_assign_kernel<<<gpts, 256, 0, s>>>(n, assignment, assignment_inv, price, bid, bid_inc, max_inc, max_idx, it == iters - 1);
  }
  emd_dist_kernel<<<gpts, 256, 0, s>>>(n, xyz1, xyz2, assignment, dist);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(3 + 4 * iters);
}

void launch_emd_backward(const float* xyz1, const float* xyz2, const float* grad_dist, const int32_t* assignment, int B, int n,
                         float* grad_xyz1, cudaStream_t s) {
  emd_grad_kernel<<<dim3((n + 255) / 256, B), 256, 0, s>>>(n, xyz1, xyz2, grad_dist, assignment, grad_xyz1);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

void launch_nn_dist(const float* a, int n, const float* b, int m, int B, int dim, float* dist, int32_t* idx, cudaStream_t s,
                    bool fma) {
  LIDM_REQUIRE(a && b && dist && idx && B > 0 && n > 0 && m > 0 && (dim == 2 || dim == 3), "nearest-neighbour distance arguments");
  dim3 grid((n + THREADS * QPT - 1) / (THREADS * QPT), B);
  if (dim == 3 && fma) nn_dist_kernel<3, true><<<grid, THREADS, 0, s>>>(a, n, b, m, dist, idx);
  else if (dim == 3) nn_dist_kernel<3, false><<<grid, THREADS, 0, s>>>(a, n, b, m, dist, idx);
  else if (fma) nn_dist_kernel<2, true><<<grid, THREADS, 0, s>>>(a, n, b, m, dist, idx);
  else nn_dist_kernel<2, false><<<grid, THREADS, 0, s>>>(a, n, b, m, dist, idx);
  LIDM_CUDA_CHECK(cudaGetLastError());
  LIDM_COUNT_LAUNCH(1);
}

}  // namespace lidm
