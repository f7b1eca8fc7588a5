"""TEST INFRASTRUCTURE ONLY (imported by tests/ only).  CPU restatement of the reference's evaluation-toolbox CUDA
extensions: the Chamfer distance (NmDistanceKernel, reference lidm/eval/modules/chamfer3D/chamfer3D.cu:12-153 and
chamfer2D/chamfer2D.cu, called by chamfer_cuda_forward :155-175; NmDistanceGradKernel :155-185) and the auction EMD
(lidm/eval/modules/emd/emd_cuda.cu:23-316).  Two forms:
  * numpy (`nn_dist`, `chamfer_forward`): fp32 with every operation rounded on its own, d = (dx*dx + dy*dy) + dz*dz;
  * plain C (oracle/eval_ref.c, compiled here with gcc and loaded through ctypes: `c_nn_dist`, `c_chamfer_backward`,
    `c_emd_forward`, `c_emd_backward`): the same with the fused multiply-adds nvcc generates for the reference sources
    written out as fmaf() - this is the form that is bit-identical to the reference extensions.

PARITY PINNED ON THE GPU: oracle/build_ref_ext.py compiles the reference extensions from /root/reference into
oracle/_ref/ (sm_100); tests/test_gpu_eval_ref.py runs them on the GPU box beside this oracle and the product (Chamfer
forward bit-exact in distances and indices; backward and EMD as far as the reference's own atomics / races are
deterministic).  Ties keep the first (lowest) index because the reference scans in ascending order with a strict `<`."""
import ctypes
import os
import subprocess

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


# ---------------------------------------------------------------------------------------------- the C form (eval_ref.c)
_HERE = os.path.dirname(os.path.abspath(__file__))
_C_SRC = os.path.join(_HERE, "eval_ref.c")
_C_LIB = os.path.join(_HERE, "_build", "liboracle_eval.so")
_clib = None


def build_c(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off: no contraction other than the fmaf() calls the source spells out."""
    if force or not os.path.exists(_C_LIB) or os.path.getmtime(_C_LIB) < os.path.getmtime(_C_SRC):
        os.makedirs(os.path.dirname(_C_LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", _C_LIB, _C_SRC, "-lm"], check=True)
    return _C_LIB


def _c():
    global _clib
    if _clib is None:
        lib = ctypes.CDLL(build_c())
        P, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        lib.oracle_nn_dist.argtypes = [P, I, P, I, I, I, I, P, P]
        lib.oracle_nn_dist.restype = None
        lib.oracle_chamfer_grad.argtypes = [P, I, P, I, I, I, P, P, P, P]
        lib.oracle_chamfer_grad.restype = None
        lib.oracle_emd_forward.argtypes = [P, P, I, I, F, I, P, P]
        lib.oracle_emd_forward.restype = I
        lib.oracle_emd_backward.argtypes = [P, P, P, P, I, I, P]
        lib.oracle_emd_backward.restype = None
        _clib = lib
    return _clib


def _f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def _i32(x):
    return np.ascontiguousarray(x, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def c_nn_dist(a, b, fma=True):
    """a (B,N,d), b (B,M,d) -> (dist (B,N) f32, idx (B,N) i32); fma=True is the reference extension's rounding."""
    a, b = _f32(a), _f32(b)
    B, N, d = a.shape
    dist, idx = np.empty((B, N), np.float32), np.empty((B, N), np.int32)
    _c().oracle_nn_dist(_p(a), N, _p(b), b.shape[1], B, d, int(fma), _p(dist), _p(idx))
    return dist, idx


def c_chamfer_forward(xyz1, xyz2, fma=True):
    d1, i1 = c_nn_dist(xyz1, xyz2, fma)
    d2, i2 = c_nn_dist(xyz2, xyz1, fma)
    return d1, d2, i1, i2


def c_chamfer_backward(xyz1, xyz2, graddist1, graddist2, idx1, idx2):
    """chamfer_cuda_backward (chamfer3D.cu:173-185): both directions accumulate into zeroed gradxyz1 / gradxyz2."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    B, N, d = xyz1.shape
    M = xyz2.shape[1]
    g1, g2 = np.zeros_like(xyz1), np.zeros_like(xyz2)
    gd1, gd2, idx1, idx2 = _f32(graddist1), _f32(graddist2), _i32(idx1), _i32(idx2)
    _c().oracle_chamfer_grad(_p(xyz1), N, _p(xyz2), M, B, d, _p(gd1), _p(idx1), _p(g1), _p(g2))
    _c().oracle_chamfer_grad(_p(xyz2), M, _p(xyz1), N, B, d, _p(gd2), _p(idx2), _p(g2), _p(g1))
    return g1, g2


def c_emd_forward(xyz1, xyz2, eps, iters):
    """emdFunction.forward (emd_module.py:47-76): -> (dist (B,n) f32 squared distances, assignment (B,n) i32)."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    B, n, _ = xyz1.shape
    if xyz2.shape != xyz1.shape:
        raise ValueError("emd: the two point clouds should have the same size")
    dist, ass = np.empty((B, n), np.float32), np.empty((B, n), np.int32)
    if _c().oracle_emd_forward(_p(xyz1), _p(xyz2), B, n, float(eps), int(iters), _p(dist), _p(ass)) != 0:
        raise ValueError("emd: batch <= 512 and n a multiple of 1024 (emd_cuda.cu:232-245)")
    return dist, ass


def c_emd_backward(xyz1, xyz2, graddist, assignment):
    xyz1, xyz2, graddist, assignment = _f32(xyz1), _f32(xyz2), _f32(graddist), _i32(assignment)
    g = np.empty_like(xyz1)
    _c().oracle_emd_backward(_p(xyz1), _p(xyz2), _p(graddist), _p(assignment), xyz1.shape[0], xyz1.shape[1], _p(g))
    return g


def compute_pairwise_emd(x, y):
    """reference lidm/eval/metric_utils.py:447-458."""
    n = min(x.shape[0], y.shape[0])
    n -= n % 1024
    dist, _ = c_emd_forward(x[None, :n], y[None, :n], 0.005, 50)
    return float(np.sqrt(dist).mean())
