"""CPU: the Chamfer nearest-neighbour oracle (oracle/eval_ref.py) against hand-computed answers and an independent
float64 brute force.  (The reference op is a CUDA extension and ships no vectors: parity for this op is unpinned.)"""
import numpy as np

from oracle import eval_ref as E


def test_known_answers_3d():
    a = np.array([[[0, 0, 0], [1, 0, 0], [5, 5, 5]]], np.float32)
    b = np.array([[[0, 0, 1], [2, 0, 0], [1, 0, 0.5], [0, 0, -1]]], np.float32)
    d1, d2, i1, i2 = E.chamfer_forward(a, b)
    # a0: |b0|^2 = 1 and |b3|^2 = 1 tie -> lowest index 0; a1: b2 at 0.25; a2: b1 at 9+25+25 = 59, b2 at 16+25+20.25
    np.testing.assert_array_equal(i1, [[0, 2, 1]])
    np.testing.assert_allclose(d1, [[1.0, 0.25, 59.0]])
    np.testing.assert_array_equal(i2, [[0, 1, 1, 0]])
    np.testing.assert_allclose(d2, [[1.0, 1.0, 0.25, 1.0]])


def test_known_answers_2d_and_ties():
    a = np.array([[[0, 0], [3, 4]]], np.float32)
    b = np.array([[[3, 4], [0, 0], [0, 0]]], np.float32)
    d1, d2, i1, i2 = E.chamfer_forward(a, b)
    np.testing.assert_array_equal(i1, [[1, 0]])          # duplicate target points: first one wins
    np.testing.assert_array_equal(d1, [[0.0, 0.0]])
    np.testing.assert_array_equal(i2, [[1, 0, 0]])
    assert abs(E.compute_pairwise_cd(a[0], b[0])) == 0.0


def test_against_float64_brute_force():
    rng = np.random.default_rng(0)
    for d in (2, 3):
        a = rng.normal(size=(2, 257, d)).astype(np.float32) * 20
        b = rng.normal(size=(2, 1031, d)).astype(np.float32) * 20
        d1, d2, i1, i2 = E.chamfer_forward(a, b)
        for x, y, dist, idx in ((a, b, d1, i1), (b, a, d2, i2)):
            dd = ((y[:, None, :, :].astype(np.float64) - x[:, :, None, :].astype(np.float64)) ** 2).sum(-1)
            np.testing.assert_allclose(dist, dd.min(-1), rtol=1e-6)
            picked = np.take_along_axis(dd, idx[..., None].astype(np.int64), -1)[..., 0]
            np.testing.assert_allclose(picked, dd.min(-1), rtol=1e-6)   # the chosen neighbour is a (near-)minimiser
        assert i1.dtype == np.int32 and d1.dtype == np.float32


def test_batch_padding_never_matches_real_points():
    rng = np.random.default_rng(1)
    ref = (rng.normal(size=(50, 3)) * 5).astype(np.float32)
    samples = [(rng.normal(size=(n, 3)) * 5).astype(np.float32) for n in (10, 50, 80)]
    batch = E.compute_pairwise_cd_batch(ref, samples)
    single = [E.compute_pairwise_cd(ref, s) for s in samples]        # un-padded, one pair at a time
    np.testing.assert_allclose(batch, single, rtol=1e-5)
