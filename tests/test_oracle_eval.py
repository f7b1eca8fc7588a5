"""CPU: the evaluation-toolbox oracle (oracle/eval_ref.py: numpy form + the plain-C form of oracle/eval_ref.c) against
hand-computed answers, an independent float64 brute force, autograd and scipy's exact assignment.  (The reference ops are
CUDA extensions: the oracle is pinned against them on the GPU box, tests/test_gpu_eval_ref.py.)"""
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


# ------------------------------------------------------------------------------------------ the C form (oracle/eval_ref.c)
def test_c_form_matches_numpy_form_without_fma():
    rng = np.random.default_rng(2)
    for d in (2, 3):
        a = (rng.normal(size=(2, 300, d)) * 20).astype(np.float32)
        b = (rng.normal(size=(2, 777, d)) * 20).astype(np.float32)
        d1, i1 = E.c_nn_dist(a, b, fma=False)
        d2, i2 = E.nn_dist(a, b)
        np.testing.assert_array_equal(d1, d2)
        np.testing.assert_array_equal(i1, i2)
        d3, i3 = E.c_nn_dist(a, b, fma=True)               # the contracted form moves a distance by at most one rounding
        np.testing.assert_allclose(d3, d2, rtol=3e-7)


def test_c_chamfer_backward_matches_autograd():
    import torch
    rng = np.random.default_rng(3)
    for d in (2, 3):
        a = (rng.normal(size=(2, 200, d)) * 5).astype(np.float32)
        b = (rng.normal(size=(2, 333, d)) * 5).astype(np.float32)
        _, _, i1, i2 = E.c_chamfer_forward(a, b)
        g1, g2 = rng.normal(size=(2, 200)).astype(np.float32), rng.normal(size=(2, 333)).astype(np.float32)
        ga, gb = E.c_chamfer_backward(a, b, g1, g2, i1, i2)
        ta, tb = torch.tensor(a, dtype=torch.float64, requires_grad=True), torch.tensor(b, dtype=torch.float64, requires_grad=True)
        n1 = torch.gather(tb, 1, torch.tensor(i1).long()[..., None].expand(-1, -1, d))
        n2 = torch.gather(ta, 1, torch.tensor(i2).long()[..., None].expand(-1, -1, d))
        loss = (((ta - n1) ** 2).sum(-1) * torch.tensor(g1)).sum() + (((tb - n2) ** 2).sum(-1) * torch.tensor(g2)).sum()
        loss.backward()
        np.testing.assert_allclose(ga, ta.grad.numpy(), rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(gb, tb.grad.numpy(), rtol=1e-4, atol=1e-4)


def test_c_emd_auction_properties():
    rng = np.random.default_rng(4)
    a, b = rng.random((2, 1024, 3), dtype=np.float32), rng.random((2, 1024, 3), dtype=np.float32)
    dist, ass = E.c_emd_forward(a, b, 0.005, 50)
    assert ass.dtype == np.int32 and ass.min() >= 0 and ass.max() < 1024
    picked = np.take_along_axis(b, ass[..., None].astype(np.int64), 1)
    np.testing.assert_allclose(dist, ((a - picked) ** 2).sum(-1), rtol=1e-5, atol=1e-7)
    # identical clouds: every point takes an object at distance ~0 (the auction may permute exact duplicates only)
    d0, a0 = E.c_emd_forward(a, a, 0.005, 50)
    assert float(d0.max()) == 0.0
    # run to convergence the auction is a permutation whose cost is within n*eps of the optimum (Bertsekas); check it
    # against scipy's exact assignment on a small eps
    from scipy.optimize import linear_sum_assignment
    dist, ass = E.c_emd_forward(a[:1], b[:1], 0.0005, 50000)
    assert len(set(ass[0].tolist())) == 1024
    cost = np.sqrt(((a[0][:, None] - b[0][None]) ** 2).sum(-1))
    r, c = linear_sum_assignment(cost)
    opt = cost[r, c].sum()
    got = np.sqrt(dist[0]).sum()
    assert opt - 1e-3 <= got <= opt + 1024 * 0.0005 + 1e-3
    g = E.c_emd_backward(a[:1], b[:1], np.ones((1, 1024), np.float32), ass)
    np.testing.assert_allclose(g, 2 * (a[:1] - np.take_along_axis(b[:1], ass[..., None].astype(np.int64), 1)), rtol=1e-6)


def test_c_emd_rejects_the_sizes_the_reference_rejects():
    import pytest
    z = np.zeros((1, 1000, 3), np.float32)
    with pytest.raises(ValueError):
        E.c_emd_forward(z, z, 0.005, 5)
    with pytest.raises(ValueError):
        E.c_emd_forward(z, np.zeros((1, 1024, 3), np.float32), 0.005, 5)
