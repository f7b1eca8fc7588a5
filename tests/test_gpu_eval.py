"""GPU (-m gpu): the Chamfer nearest-neighbour kernel through the C ABI (`lidm_chamfer_nn`, mirrored by
`eval_ops.chamfer_3DDist` / `chamfer_2DDist`) against the oracle: distances and indices bit-exact (same fp32 operation
order, ties to the lowest index), ragged sizes, padded clouds as `compute_pairwise_cd_batch` builds them, and
size-independent properties at range-image scale."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import eval_ref as E


@pytest.mark.parametrize("dim", [3, 2])
@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 257, 1031), (3, 1024, 1024), (1, 2500, 513)])
def test_matches_oracle_bit_exact(built_lib, dim, shape):
    from lidar_layout_b200.eval_ops import chamfer_2DDist, chamfer_3DDist
    B, N, M = shape
    rng = np.random.default_rng(B * 1000 + N + dim)
    a = (rng.normal(size=(B, N, dim)) * 20).astype(np.float32)
    b = (rng.normal(size=(B, M, dim)) * 20).astype(np.float32)
    b[:, M // 2] = b[:, 0]                                     # duplicated target point: the lower index must win
    mod = chamfer_3DDist() if dim == 3 else chamfer_2DDist()
    d1, d2, i1, i2 = mod(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    r1, r2, j1, j2 = E.chamfer_forward(a, b)
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
