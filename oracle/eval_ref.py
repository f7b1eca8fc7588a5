"""TEST INFRASTRUCTURE ONLY (imported by tests/ only).  CPU restatement of the forward of the reference's Chamfer-distance
extension: NmDistanceKernel (reference lidm/eval/modules/chamfer3D/chamfer3D.cu:12-155, chamfer2D/chamfer2D.cu:12-145)
as called by chamfer_cuda_forward (chamfer3D.cu:155-175) - for every point of one set the squared distance to, and the
index of, its nearest point in the other set; ties keep the first (lowest) index because the kernel scans in
ascending order with a strict `<`.

PARITY UNPINNED: the reference implementation is a CUDA extension, so it cannot run in the build container (no GPU), and
/root/reference does not exist on the GPU box; it ships no test vectors for this op.  The restatement is pinned only by
hand-computed known answers and by an independent float64 brute force (tests/test_oracle_eval.py).  fp32 arithmetic in
the order d = (dx*dx + dy*dy) + dz*dz without fused multiply-add; the reference's own build may contract these into
FMAs, which moves a distance by at most one rounding and can only swap indices between points that are equidistant to
within that rounding."""
import numpy as np


def nn_dist(a: np.ndarray, b: np.ndarray):
    """a (B,N,d), b (B,M,d) float32 -> (dist (B,N) float32, idx (B,N) int32)."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    B, N, d = a.shape
    dist = np.empty((B, N), np.float32)
    idx = np.empty((B, N), np.int32)
    for i in range(B):
        dx = b[i, None, :, 0] - a[i, :, None, 0]
        dy = b[i, None, :, 1] - a[i, :, None, 1]
        dd = dx * dx + dy * dy                      # float32 throughout: each product and sum rounded on its own
        if d == 3:
            dz = b[i, None, :, 2] - a[i, :, None, 2]
            dd = dd + dz * dz
        j = np.argmin(dd, axis=1)                   # first minimum = lowest index
        idx[i] = j.astype(np.int32)
        dist[i] = dd[np.arange(N), j]
    return dist, idx


def chamfer_forward(xyz1: np.ndarray, xyz2: np.ndarray):
    """chamfer_3DFunction.forward / chamfer_2DFunction.forward: -> dist1, dist2, idx1, idx2."""
    d1, i1 = nn_dist(xyz1, xyz2)
    d2, i2 = nn_dist(xyz2, xyz1)
    return d1, d2, i1, i2


def compute_pairwise_cd(x: np.ndarray, y: np.ndarray) -> float:
    """reference lidm/eval/metric_utils.py:414-423."""
    if x.ndim == 2:
        x, y = x[None], y[None]
    d1, d2, _, _ = chamfer_forward(x, y)
    return float((np.float32(d1.mean(dtype=np.float32)) + np.float32(d2.mean(dtype=np.float32))) / 2)


def compute_pairwise_cd_batch(reference: np.ndarray, samples):
    """reference lidm/eval/metric_utils.py:426-444 (padding with 1e6 points, means over the un-padded prefix)."""
    d = reference.shape[-1]
    len_r, len_s = reference.shape[0], [s.shape[0] for s in samples]
    width = max([len_r] + len_s)
    pad = lambda c: np.vstack([c.astype(np.float32), np.full((width - c.shape[0], d), 1e6, np.float32)])
    smp = np.stack([pad(c) for c in samples])
    ref = np.broadcast_to(pad(reference), smp.shape)
    d_r, d_s, _, _ = chamfer_forward(ref, smp)
    return [float((d_r[i, :len_r].mean(dtype=np.float32) + d_s[i, :n].mean(dtype=np.float32)) / np.float32(2)) for i, n in enumerate(len_s)]
