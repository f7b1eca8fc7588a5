"""GPU (-m gpu): the product's evaluation kernels and the C oracle against THE REFERENCE'S OWN CUDA EXTENSIONS - chamfer3D and emd,
compiled unmodified from /root/reference into oracle/_ref/ by oracle/build_ref_ext.py (the .so files travel to the GPU box; the
sources do not).  This is what pins oracle/eval_ref.c:
  * Chamfer forward: distances and indices bit-identical (which settles the FMA contraction: nvcc fuses the reference's
    `dx*dx + dy*dy + dz*dz`, and so does the product by default);
  * Chamfer backward: equal up to the order of the reference's atomicAdds;
  * EMD: the reference's GetMax lets the last of several near-equal top bidders win (a data race), so the comparison is exact
    wherever the run had no such tie and statistical otherwise: assignments agree on almost every point and the EMD value
    agrees to 1e-3 relative."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import build_ref_ext as RX
from oracle import eval_ref as E


@pytest.fixture(scope="module")
def ref_chamfer():
    mod = RX.load_ext("ref_chamfer_3D")
    if mod is None:
        pytest.skip("oracle/_ref/ref_chamfer_3D not built (python -m oracle.build_ref_ext, needs /root/reference)")
    return mod


@pytest.fixture(scope="module")
def ref_emd():
    mod = RX.load_ext("ref_emd")
    if mod is None:
        pytest.skip("oracle/_ref/ref_emd not built (python -m oracle.build_ref_ext, needs /root/reference)")
    return mod


def _ref_chamfer_forward(ext, a, b):
    """chamfer_3DFunction.forward (dist_chamfer_3D.py:30-52)."""
    B, N, _ = a.shape
    M = b.shape[1]
    d1, d2 = torch.zeros(B, N, device="cuda"), torch.zeros(B, M, device="cuda")
    i1, i2 = torch.zeros(B, N, device="cuda", dtype=torch.int32), torch.zeros(B, M, device="cuda", dtype=torch.int32)
    ext.forward(a, b, d1, d2, i1, i2)
    return d1, d2, i1, i2


def _ref_emd_forward(ext, a, b, eps, iters):
    """emdFunction.forward (emd_module.py:47-76): the workspaces exactly as the reference allocates them."""
    B, n, _ = a.shape
    z = lambda *s, dt=torch.float32: torch.zeros(*s, device="cuda", dtype=dt)
    dist, assignment, assignment_inv = z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n, dt=torch.int32) - 1
    price, bid, bid_inc, max_inc = z(B, n), z(B, n, dt=torch.int32), z(B, n), z(B, n)
    unass_idx, max_idx = z(B * n, dt=torch.int32), z(B * n, dt=torch.int32)
    unass_cnt, unass_cnt_sum, cnt_tmp = z(512, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32)
    ext.forward(a, b, dist, assignment, price, assignment_inv, bid, bid_inc, max_inc, unass_idx, unass_cnt, unass_cnt_sum,
                cnt_tmp, max_idx, eps, iters)
    torch.cuda.synchronize()
    return dist, assignment


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 257, 1031), (3, 1024, 1024), (2, 8192, 5000)])
def test_chamfer_forward_bit_identical_to_the_reference_extension(built_lib, ref_chamfer, shape):
    from lidar_layout_b200.eval_ops import chamfer_3DDist
    B, N, M = shape
    g = torch.Generator(device="cuda").manual_seed(N)
    a = torch.randn(B, N, 3, device="cuda", generator=g) * 20
    b = torch.randn(B, M, 3, device="cuda", generator=g) * 20
    b[:, M // 2] = b[:, 0]
    r1, r2, j1, j2 = _ref_chamfer_forward(ref_chamfer, a, b)
    d1, d2, i1, i2 = chamfer_3DDist()(a, b)
    assert torch.equal(i1, j1) and torch.equal(i2, j2)
    assert torch.equal(d1, r1) and torch.equal(d2, r2)
    if N * M <= 1 << 21:                                         # the C oracle too (sequential: small cases only)
        o1, o2, k1, k2 = E.c_chamfer_forward(a.cpu().numpy(), b.cpu().numpy(), fma=True)
        np.testing.assert_array_equal(o1, r1.cpu().numpy())
        np.testing.assert_array_equal(k1, j1.cpu().numpy())
        np.testing.assert_array_equal(o2, r2.cpu().numpy())
        np.testing.assert_array_equal(k2, j2.cpu().numpy())


def test_chamfer_backward_against_the_reference_extension(built_lib, ref_chamfer):
    from lidar_layout_b200.eval_ops import chamfer_3DDist
    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(2, 3000, 3, device="cuda", generator=g) * 10).requires_grad_()
    b = (torch.randn(2, 2000, 3, device="cuda", generator=g) * 10).requires_grad_()
    g1, g2 = torch.randn(2, 3000, device="cuda", generator=g), torch.randn(2, 2000, device="cuda", generator=g)
    d1, d2, i1, i2 = chamfer_3DDist()(a, b)
    ((d1 * g1).sum() + (d2 * g2).sum()).backward()
    ra, rb = torch.zeros_like(a), torch.zeros_like(b)
    ref_chamfer.backward(a.detach(), b.detach(), ra, rb, g1, g2, i1, i2)
    assert torch.allclose(a.grad, ra, rtol=1e-5, atol=1e-4) and torch.allclose(b.grad, rb, rtol=1e-5, atol=1e-4)
    oa, ob = E.c_chamfer_backward(a.detach().cpu().numpy(), b.detach().cpu().numpy(), g1.cpu().numpy(), g2.cpu().numpy(),
                                  i1.cpu().numpy(), i2.cpu().numpy())
    np.testing.assert_allclose(oa, ra.cpu().numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(ob, rb.cpu().numpy(), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("shape,eps,iters", [((2, 1024), 0.005, 50), ((4, 2048), 0.005, 50), ((1, 8192), 0.005, 50), ((2, 1024), 0.005, 1)])
def test_emd_against_the_reference_extension(built_lib, ref_emd, shape, eps, iters):
    from lidar_layout_b200.eval_ops import emdModule
    B, n = shape
    g = torch.Generator(device="cuda").manual_seed(n + B)
    a, b = torch.rand(B, n, 3, device="cuda", generator=g), torch.rand(B, n, 3, device="cuda", generator=g)
    rd, ra = _ref_emd_forward(ref_emd, a, b, eps, iters)
    dist, ass = emdModule()(a, b, eps, iters)
    agree = (ass == ra).float().mean().item()
    emd_ref, emd_got = rd.sqrt().mean(1), dist.sqrt().mean(1)
    print(f"emd B={B} n={n} iters={iters}: assignment agreement {agree:.4f}, EMD ref {emd_ref.tolist()} got {emd_got.tolist()}")
    if iters == 1:                                               # one round: no price has moved, no eviction - no race to lose
        assert agree == 1.0 and torch.equal(dist, rd)
    assert agree > 0.9
    assert torch.allclose(emd_got, emd_ref, rtol=1e-3)
    # wherever both chose the same object the distance is the same bits (CalcDist, emd_cuda.cu:211-221)
    same = ass == ra
    assert torch.equal(dist[same], rd[same])


def test_emd_backward_against_the_reference_extension(built_lib, ref_emd):
    from lidar_layout_b200.eval_ops import emdModule
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.rand(2, 2048, 3, device="cuda", generator=g).requires_grad_()
    b = torch.rand(2, 2048, 3, device="cuda", generator=g)
    gd = torch.randn(2, 2048, device="cuda", generator=g)
    dist, ass = emdModule()(a, b, 0.005, 50)
    (dist * gd).sum().backward()
    ref = torch.zeros_like(a)
    ref_emd.backward(a.detach(), b, ref, gd, ass)
    torch.cuda.synchronize()
    assert torch.equal(a.grad, ref)
