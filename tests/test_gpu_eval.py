"""GPU (-m gpu): the evaluation-toolbox kernels through the C ABI (`lidm_chamfer_nn[_ex]`, `lidm_chamfer_backward`,
`lidm_emd_forward`, `lidm_emd_backward`, mirrored by `eval_ops.chamfer_3DDist` / `chamfer_2DDist` / `emdModule`) against the
oracle: distances, indices and auction assignments bit-exact (same fp32 operation order, ties to the lowest index), ragged
sizes, padded clouds as `compute_pairwise_cd_batch` builds them, and size-independent properties at range-image scale.
tests/test_gpu_eval_ref.py pins the same against the reference's own extensions."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import eval_ref as E


@pytest.mark.parametrize("fma", [True, False])
@pytest.mark.parametrize("dim", [3, 2])
@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 257, 1031), (3, 1024, 1024), (1, 2500, 513)])
def test_matches_oracle_bit_exact(built_lib, dim, shape, fma):
    from lidar_layout_b200.eval_ops import chamfer_2DDist, chamfer_3DDist
    B, N, M = shape
    rng = np.random.default_rng(B * 1000 + N + dim)
    a = (rng.normal(size=(B, N, dim)) * 20).astype(np.float32)
    b = (rng.normal(size=(B, M, dim)) * 20).astype(np.float32)
    b[:, M // 2] = b[:, 0]                                     # duplicated target point: the lower index must win
    mod = chamfer_3DDist() if dim == 3 else chamfer_2DDist()
    mod.CONTRACT_FMA = fma
    d1, d2, i1, i2 = mod(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    # fma: the reference extension's rounding (the C oracle spells the fmaf() out); else the numpy form
    r1, r2, j1, j2 = E.c_chamfer_forward(a, b, fma=True) if fma else E.chamfer_forward(a, b)
    assert d1.dtype == torch.float32 and i1.dtype == torch.int32 and d1.shape == (B, N) and i2.shape == (B, M)
    np.testing.assert_array_equal(i1.cpu().numpy(), j1)
    np.testing.assert_array_equal(i2.cpu().numpy(), j2)
    np.testing.assert_array_equal(d1.cpu().numpy(), r1)
    np.testing.assert_array_equal(d2.cpu().numpy(), r2)


def test_padded_clouds_and_pairwise_cd(built_lib):
    from lidar_layout_b200.eval_ops import compute_pairwise_cd
    rng = np.random.default_rng(3)
    x = (rng.normal(size=(700, 3)) * 10).astype(np.float32)
    y = (rng.normal(size=(900, 3)) * 10).astype(np.float32)
    assert compute_pairwise_cd(x, y) == pytest.approx(E.compute_pairwise_cd(x, y), rel=1e-6)
    # the reference pads ragged clouds with points at 1e6 (metric_utils.py:431-436): they must only ever match each other
    xp = np.vstack([x, np.full((200, 3), 1e6, np.float32)])
    assert compute_pairwise_cd(xp, y) == pytest.approx(E.compute_pairwise_cd(xp, y), rel=1e-6)


def test_pairwise_cd_batch_ragged(built_lib):
    from lidar_layout_b200.eval_ops import compute_pairwise_cd_batch
    rng = np.random.default_rng(5)
    for d in (3, 2):
        ref = (rng.normal(size=(600, d)) * 10).astype(np.float32)
        samples = [(rng.normal(size=(n, d)) * 10).astype(np.float32) for n in (1, 333, 600, 901)]
        got = compute_pairwise_cd_batch(ref, samples)
        want = E.compute_pairwise_cd_batch(ref, samples)
        assert len(got) == 4 and got == pytest.approx(want, rel=1e-5)


def test_properties_at_range_image_scale(built_lib):
    from lidar_layout_b200.eval_ops import chamfer_3DDist
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(2, 65536, 3, device="cuda", generator=g) * 30
    b = torch.randn(2, 50000, 3, device="cuda", generator=g) * 30
    mod = chamfer_3DDist()
    d1, d2, i1, i2 = mod(a, b)
    e2, e1, k2, k1 = mod(b, a)                                  # swapping the sets swaps the outputs
    assert torch.equal(d1, e1) and torch.equal(i1, k1) and torch.equal(d2, e2) and torch.equal(i2, k2)
    assert int(i1.min()) >= 0 and int(i1.max()) < b.shape[1] and float(d1.min()) >= 0
    near = torch.gather(b, 1, i1.long()[..., None].expand(-1, -1, 3))
    assert torch.allclose(((near - a) ** 2).sum(-1), d1, rtol=1e-5, atol=1e-6)      # the index really is that point
    s0, _, j0, _ = mod(a, a)                                    # a set against itself: distance 0 at its own index
    assert float(s0.max()) == 0.0 and torch.equal(j0, torch.arange(65536, device="cuda", dtype=torch.int32).expand(2, -1))
    sub = torch.randperm(65536, device="cuda", generator=g)[:4096]
    ref = torch.cdist(a[:, sub].double(), b.double()).min(-1).values ** 2
    assert torch.allclose(d1[:, sub].double(), ref, rtol=1e-5, atol=1e-6)


def test_error_behaviour(built_lib):
    from lidar_layout_b200.eval_ops import chamfer_3DDist
    with pytest.raises(ValueError):
        chamfer_3DDist()(torch.zeros(1, 4, 2, device="cuda"), torch.zeros(1, 4, 3, device="cuda"))
    with pytest.raises(ValueError):
        chamfer_3DDist()(torch.zeros(1, 4, 3), torch.zeros(1, 4, 3))


@pytest.mark.parametrize("dim", [3, 2])
def test_chamfer_backward_against_oracle(built_lib, dim):
    from lidar_layout_b200.eval_ops import chamfer_2DDist, chamfer_3DDist
    rng = np.random.default_rng(11 + dim)
    a = (rng.normal(size=(3, 700, dim)) * 10).astype(np.float32)
    b = (rng.normal(size=(3, 1300, dim)) * 10).astype(np.float32)
    g1, g2 = rng.normal(size=(3, 700)).astype(np.float32), rng.normal(size=(3, 1300)).astype(np.float32)
    ta, tb = torch.from_numpy(a).cuda().requires_grad_(), torch.from_numpy(b).cuda().requires_grad_()
    d1, d2, i1, i2 = (chamfer_3DDist() if dim == 3 else chamfer_2DDist())(ta, tb)
    ((d1 * torch.from_numpy(g1).cuda()).sum() + (d2 * torch.from_numpy(g2).cuda()).sum()).backward()
    ra, rb = E.c_chamfer_backward(a, b, g1, g2, i1.cpu().numpy(), i2.cpu().numpy())
    # the scatter into the other set is a sum of atomics (order free, as in the reference): equal up to reassociation
    np.testing.assert_allclose(ta.grad.cpu().numpy(), ra, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(tb.grad.cpu().numpy(), rb, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("shape,eps,iters", [((2, 1024), 0.005, 50), ((3, 2048), 0.005, 50), ((1, 4096), 0.002, 20), ((1, 1024), 0.005, 1)])
def test_emd_matches_oracle_bit_exact(built_lib, shape, eps, iters):
    """The auction through the C ABI (lidm_emd_forward) against the sequential C restatement: assignment and distances
    identical (same arithmetic, same tie rules), gradient identical."""
    from lidar_layout_b200.eval_ops import emdModule
    B, n = shape
    rng = np.random.default_rng(B * 7 + n)
    a, b = rng.random((B, n, 3), dtype=np.float32), rng.random((B, n, 3), dtype=np.float32)
    b[:, 5] = b[:, 900]                                      # duplicated object: equal values go to the lower index
    ta = torch.from_numpy(a).cuda().requires_grad_()
    dist, ass = emdModule()(ta, torch.from_numpy(b).cuda(), eps, iters)
    rd, ra = E.c_emd_forward(a, b, eps, iters)
    assert dist.dtype == torch.float32 and ass.dtype == torch.int32 and dist.shape == (B, n)
    np.testing.assert_array_equal(ass.cpu().numpy(), ra)
    np.testing.assert_array_equal(dist.detach().cpu().numpy(), rd)
    g = rng.normal(size=(B, n)).astype(np.float32)
    (dist * torch.from_numpy(g).cuda()).sum().backward()
    np.testing.assert_array_equal(ta.grad.cpu().numpy(), E.c_emd_backward(a, b, g, ra))


def test_emd_pairwise_and_limits(built_lib):
    from lidar_layout_b200.eval_ops import compute_pairwise_emd, emdModule
    rng = np.random.default_rng(9)
    x, y = rng.random((2500, 3), dtype=np.float32), rng.random((2100, 3), dtype=np.float32)   # truncated to 2048 points
    assert compute_pairwise_emd(x, y) == pytest.approx(E.compute_pairwise_emd(x, y), rel=1e-6)
    with pytest.raises(RuntimeError, match="multiple of 1024"):                              # emd_cuda.cu:242-245
        emdModule()(torch.zeros(1, 1000, 3).cuda(), torch.zeros(1, 1000, 3).cuda(), 0.005, 5)
    with pytest.raises(AssertionError):                                                      # emd_module.py:53
        emdModule()(torch.zeros(1, 1024, 3).cuda(), torch.zeros(1, 2048, 3).cuda(), 0.005, 5)
