/* TEST INFRASTRUCTURE ONLY (loaded by oracle/eval_ref.py for tests/ only; never linked into the product).
 *
 * Plain-C restatement of the arithmetic of the reference's evaluation-toolbox CUDA extensions, sequential and
 * deterministic, with the fused multiply-adds nvcc generates for the reference sources written out as fmaf():
 *   - NmDistanceKernel      lidm/eval/modules/chamfer3D/chamfer3D.cu:12-153 (and chamfer2D): nearest neighbour + index
 *   - NmDistanceGradKernel  lidm/eval/modules/chamfer3D/chamfer3D.cu:155-171: Chamfer backward
 *   - the auction EMD       lidm/eval/modules/emd/emd_cuda.cu:23-284 (Bid / GetMax / Assign / CalcDist) and :286-303 (grad)
 * Pinned on the GPU against the reference extensions themselves, built from /root/reference by oracle/build_ref_ext.py
 * (tests/test_gpu_eval_ref.py).  Build: gcc -O2 -ffp-contract=off -shared -fPIC (oracle/eval_ref.py does it). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float sqdist(const float* p, const float* q, int dim, int fma) {
  /* `dx*dx + dy*dy + dz*dz` as nvcc contracts it (SASS of the reference extension built for sm_100: FMUL dy*dy, FFMA dx*dx + .,
   * FFMA dz*dz + .): fma(dz, dz, fma(dx, dx, dy*dy)); fma == 0: every operation rounded on its own */
  const float dx = q[0] - p[0], dy = q[1] - p[1];
  float d = fma ? fmaf(dx, dx, dy * dy) : dx * dx + dy * dy;
  if (dim == 3) {
    const float dz = q[2] - p[2];
    d = fma ? fmaf(dz, dz, d) : d + dz * dz;
  }
  return d;
}

/* a (B,n,dim), b (B,m,dim) -> dist (B,n), idx (B,n): ascending scan, strict <, so ties keep the lowest index */
void oracle_nn_dist(const float* a, int n, const float* b, int m, int B, int dim, int fma, float* dist, int32_t* idx) {
  for (int i = 0; i < B; ++i)
    for (int j = 0; j < n; ++j) {
      const float* p = a + ((size_t)i * n + j) * dim;
      float best = 0;
      int best_i = 0;
      for (int k = 0; k < m; ++k) {
        const float d = sqdist(p, b + ((size_t)i * m + k) * dim, dim, fma);
        if (k == 0 || d < best) { best = d; best_i = k; }
      }
      dist[(size_t)i * n + j] = best;
      idx[(size_t)i * n + j] = best_i;
    }
}

/* one direction of the Chamfer backward; grad_a (B,n,dim) and grad_b (B,m,dim) accumulate (callers zero them) */
void oracle_chamfer_grad(const float* a, int n, const float* b, int m, int B, int dim, const float* grad_dist, const int32_t* idx,
                         float* grad_a, float* grad_b) {
  for (int i = 0; i < B; ++i)
    for (int j = 0; j < n; ++j) {
      const int j2 = idx[(size_t)i * n + j];
      const float g = grad_dist[(size_t)i * n + j] * 2;
      for (int c = 0; c < dim; ++c) {
        const float v = g * (a[((size_t)i * n + j) * dim + c] - b[((size_t)i * m + j2) * dim + c]);
        grad_a[((size_t)i * n + j) * dim + c] += v;
        grad_b[((size_t)i * m + j2) * dim + c] += -v;
      }
    }
}

/* The auction.  Where the reference leaves the outcome to a data race (several bidders of one object within 1e-6 of the
 * highest increment: the last writer of max_idx wins, emd_cuda.cu:184-187) the bidder with the highest point index wins.
 * Returns 0, or -1 for the sizes the reference rejects (emd_cuda.cu:232-245). */
int oracle_emd_forward(const float* xyz1, const float* xyz2, int B, int n, float eps, int iters, float* dist, int32_t* assignment) {
  if (B > 512 || n % 1024 != 0) return -1;
  int32_t* inv = malloc(sizeof(int32_t) * n);
  int32_t* bid = calloc(n, sizeof(int32_t));
  int32_t* max_idx = malloc(sizeof(int32_t) * n);
  uint8_t* open = malloc(n);
  float* price = malloc(sizeof(float) * n);
  float* bid_inc = calloc(n, sizeof(float));
  float* max_inc = malloc(sizeof(float) * n);
  for (int i = 0; i < B; ++i) {
    const float* p1 = xyz1 + (size_t)i * n * 3;
    const float* p2 = xyz2 + (size_t)i * n * 3;
    int32_t* ass = assignment + (size_t)i * n;
    for (int j = 0; j < n; ++j) { ass[j] = -1; inv[j] = -1; price[j] = 0; max_inc[j] = 0; }
    for (int it = 0; it < iters; ++it) {
      const int last = it == iters - 1;
      for (int j = 0; j < n; ++j) { open[j] = ass[j] == -1; max_idx[j] = -1; }
      for (int j = 0; j < n; ++j) {                                   /* Bid */
        if (!open[j]) continue;
        float best = -1e9f, better = -1e9f;
        int best_i = -1;
        for (int k = 0; k < n; ++k) {
          const float s2 = sqdist(p1 + 3 * j, p2 + 3 * k, 3, 1);
          const float d = (float)(3.0 - (double)sqrtf(s2) - (double)price[k]);
          if (d > best) { better = best; best = d; best_i = k; }
          else if (d > better) better = d;
        }
        const float inc = best - better + eps;
        bid[j] = best_i;
        bid_inc[j] = inc;
        if (inc > max_inc[best_i]) max_inc[best_i] = inc;
      }
      for (int j = 0; j < n; ++j) {                                   /* GetMax */
        if (!open[j]) continue;
        const float bi = bid_inc[j], mi = max_inc[bid[j]];
        if (bi - 1e-6 <= mi && mi <= bi + 1e-6 && j > max_idx[bid[j]]) max_idx[bid[j]] = j;
      }
      for (int j = 0; j < n; ++j) {                                   /* Assign */
        if (!open[j]) continue;
        const int o = bid[j];
        if (last || max_idx[o] == j) {
          if (!last && inv[o] != -1) ass[inv[o]] = -1;
          inv[o] = j;
          ass[j] = o;
          price[o] += bid_inc[j];
          max_inc[o] = -1e9f;
        }
      }
    }
    for (int j = 0; j < n; ++j) {                                     /* CalcDist */
      const float* q = p2 + 3 * (size_t)ass[j];
      const float dx = p1[3 * j] - q[0], dy = p1[3 * j + 1] - q[1], dz = p1[3 * j + 2] - q[2];
      dist[(size_t)i * n + j] = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
    }
  }
  free(inv); free(bid); free(max_idx); free(open); free(price); free(bid_inc); free(max_inc);
  return 0;
}

/* emd_cuda.cu:286-303: only xyz1 receives a gradient */
void oracle_emd_backward(const float* xyz1, const float* xyz2, const float* grad_dist, const int32_t* assignment, int B, int n,
                         float* grad_xyz1) {
  for (size_t i = 0; i < (size_t)B; ++i)
    for (int j = 0; j < n; ++j) {
      const int k = assignment[i * n + j];
      const float g = grad_dist[i * n + j] * 2;
      for (int c = 0; c < 3; ++c) grad_xyz1[(i * n + j) * 3 + c] = g * (xyz1[(i * n + j) * 3 + c] - xyz2[(i * n + k) * 3 + c]);
    }
}
