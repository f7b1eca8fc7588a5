"""GPU (-m gpu): operator-level parity of the CUDA kernels, called through the C ABI, against the oracle on the
same seeded inputs.  Tolerances: integer/index work and the DDIM arithmetic bit-exact; bf16 tensor-core ops are
compared against the oracle evaluated on bf16-rounded operands (tolerance = accumulation order only, 2e-5) and
against the fp32 oracle within the bf16 budget stated in BASELINE.json north_star (2e-2)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import torch_ref as R


def rel(a, b):
    return R.rel_l2(a.detach().cpu(), b.detach().cpu() if isinstance(b, torch.Tensor) else b)


@pytest.fixture(scope="module")
def ops(built_lib):
    assert torch.cuda.is_available(), "-m gpu tests need a B200"
    from lidar_layout_b200 import ops as _ops
    return _ops


CONV_CASES = [
    # B, Cin, H, W, Cout, kh, kw, pad(l,r,t,b), stride           what it is in the model
    (1, 64, 4, 32, 128, 1, 1, (0, 0, 0, 0), 1),                 # 1x1 skip
    (2, 256, 16, 128, 256, 3, 3, (1, 1, 1, 1), 1),              # level-0 ResBlock conv
    (2, 512, 8, 64, 512, 3, 3, (1, 1, 1, 1), 1),                # level-1 (2-row tiles)
    (2, 1024, 4, 32, 1024, 3, 3, (1, 1, 1, 1), 1),              # level-2 (4-row tiles)
    (2, 768, 16, 128, 256, 3, 3, (1, 1, 1, 1), 1),              # concat input
    (2, 8, 16, 128, 256, 3, 3, (1, 1, 1, 1), 1),                # input conv (im2col, K padded)
    (2, 256, 16, 128, 8, 3, 3, (1, 1, 1, 1), 1),                # out conv (N = 8)
    (2, 256, 16, 128, 256, 3, 3, (1, 1, 1, 1), 2),              # Downsample: strided TMA boxes, 64-pixel rows
    (2, 512, 8, 64, 512, 3, 3, (1, 1, 1, 1), 2),                # Downsample to 4x32 (4-row tiles)
    (1, 64, 8, 512, 128, 3, 3, (1, 1, 1, 1), 2),                # stride 2 at 256-pixel output rows (two tiles per row)
    (1, 128, 8, 256, 128, 1, 4, (1, 2, 0, 0), 1),               # decoder curve-wise (1,4), asymmetric pad
    (1, 128, 8, 256, 128, 1, 5, (2, 2, 0, 0), 1),               # decoder Upsample conv (1,5)
    (1, 64, 8, 256, 1, 1, 4, (1, 2, 0, 0), 1),                  # decoder conv_out
    (1, 128, 8, 256, 64, 1, 4, (1, 2, 0, 0), 1),                # Cout = 64 tile
    (2, 64, 8, 256, 64, 3, 3, (1, 1, 1, 1), 1),                 # 64 -> 64 ring conv (R2DM level 0): halo-tile kernel
    (1, 64, 4, 512, 64, 1, 4, (1, 2, 0, 0), 1),                 # 64 -> 64 (1,4) (decoder full resolution): halo-tile kernel
    (1, 64, 4, 128, 64, 1, 5, (2, 2, 0, 0), 1),                 # 64 -> 64 (1,5), one tile per row
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_circular_conv2d(ops, case):
    B, Cin, H, W, Cout, kh, kw, pad, stride = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, kh, kw, generator=g) / (Cin * kh * kw) ** 0.5
    b = torch.randn(Cout, generator=g)
    y = ops.circular_conv2d(x.cuda(), w.cuda(), b.cuda(), pad, stride)
    ref_q = R.circular_conv2d(x.bfloat16().float(), w.bfloat16().float(), b, pad, stride)
    ref = R.circular_conv2d(x, w, b, pad, stride)
    assert y.shape == ref.shape
    assert rel(y, ref_q) < 2e-5
    assert rel(y, ref) < 2e-2


def test_conv_residual_and_wraparound(ops):
    g = torch.Generator().manual_seed(5)
    x = torch.zeros(1, 64, 4, 32)
    x[:, :, :, 0] = torch.randn(1, 64, 4, generator=g)        # only column 0 is non-zero
    w = torch.randn(64, 64, 3, 3, generator=g) / 24
    res = torch.randn(1, 64, 4, 32, generator=g)
    y = ops.circular_conv2d(x.cuda(), w.cuda(), None, (1, 1, 1, 1), 1, residual=res.cuda())
    ref = R.circular_conv2d(x.bfloat16().float(), w.bfloat16().float(), None, (1, 1, 1, 1)) + res.bfloat16().float()
    assert rel(y, ref) < 2e-5
    # circular: the last column must see column 0 through the wrap, zero padding must not leak across H
    assert float((y.cpu() - res.bfloat16().float())[:, :, :, 31].abs().max()) > 0


GN_CASES = [(2, 256, 16, 128), (1, 768, 16, 128), (1, 1536, 8, 64), (2, 2048, 4, 32), (2, 64, 8, 64), (2, 128, 4, 32),
            (1, 192, 8, 64)]


@pytest.mark.parametrize("shape", GN_CASES)
@pytest.mark.parametrize("silu", [False, True])
def test_groupnorm(ops, shape, silu):
    B, C, H, W = shape
    g = torch.Generator().manual_seed(C + H)
    x = torch.randn(B, C, H, W, generator=g) * 2 + 0.3
    ga, be = torch.randn(C, generator=g), torch.randn(C, generator=g)
    y = ops.group_norm(x.cuda(), ga.cuda(), be.cuda(), 1e-5, 32, silu)
    r = F.group_norm(x.bfloat16().float(), 32, ga, be, 1e-5)
    r = F.silu(r) if silu else r
    # output is stored as bf16: half-ulp rounding = 2^-9 relative per element
    assert rel(y, r) < 3e-3
    assert float((y.cpu() - r).abs().max()) < 2 ** -7 * float(r.abs().max())


@pytest.mark.parametrize("B,heads,T", [(1, 2, 128), (2, 4, 512), (1, 8, 2048), (2, 32, 128)])
def test_qkv_attention_legacy(ops, B, heads, T):
    g = torch.Generator().manual_seed(T + heads)
    qkv = torch.randn(B, heads * 96, T, generator=g)
    y = ops.qkv_attention_legacy(qkv.cuda(), heads)
    ref = R.qkv_attention_legacy(qkv, heads)
    assert rel(y, ref) < 1e-2
    # rows of softmax sum to one: attention of a constant V returns that constant
    qkv2 = qkv.clone().reshape(B * heads, 96, T)
    qkv2[:, 64:, :] = 1.25
    y2 = ops.qkv_attention_legacy(qkv2.reshape(B, heads * 96, T).cuda(), heads)
    assert float((y2 - 1.25).abs().max()) < 1e-2


def test_ddim_step_bit_exact(ops):
    g = torch.Generator().manual_seed(0)
    for n in ((2, 8, 16, 128), (3, 5, 7)):          # second shape exercises the scalar tail
        x, e, nz = (torch.randn(*n, generator=g) for _ in range(3))
        coef = np.float32([0.5123, 0.6234, 0.1, np.sqrt(np.float32(1 - 0.5123))])
        xp, x0 = ops.ddim_step(x.cuda(), e.cuda(), coef, nz.cuda(), 0.9)
        rp, r0 = R.ddim_step(x.reshape(n[0], -1, 1, 1), e.reshape(n[0], -1, 1, 1), coef,
                             nz.reshape(n[0], -1, 1, 1), 0.9)
        assert torch.equal(xp.cpu().reshape(rp.shape), rp) and torch.equal(x0.cpu().reshape(r0.shape), r0)
        xp, _ = ops.ddim_step(x.cuda(), e.cuda(), coef, None, 1.0)
        rp, _ = R.ddim_step(x.reshape(n[0], -1, 1, 1), e.reshape(n[0], -1, 1, 1), coef)
        assert torch.equal(xp.cpu().reshape(rp.shape), rp)


def test_backproject_matches_reference_fixture(ops, golden_kitti):
    g = golden_kitti
    img = torch.from_numpy(g["bp_img"]).cuda()[None]
    xyz, mask = ops.backproject(img, (3, -25), (1.0, 56.0), 5.84, True)
    ref = g["bp_xyz"]
    # fp32 kernel vs the reference's float64 geometry on float32 depth: <= 2 ulp of 56 m
    assert float(np.abs(xyz[0].cpu().numpy().astype(np.float64) - ref).max()) < 2e-5
    assert int(mask.sum()) == g["bp_pcd"].shape[0]                      # exactly the points range2pcd keeps
    pts = xyz[0].reshape(3, -1).t()[mask[0].flatten().bool()].cpu().numpy()
    assert np.abs(pts - g["bp_pcd"]).max() < 2e-5


def test_backproject_edges_and_full_size(ops):
    ds = dict(fov=(3, -25), depth_range=(1.0, 56.0), depth_scale=5.84, log_scale=True)
    empty = torch.full((2, 64, 1024), -1.0).cuda()
    xyz, mask = ops.backproject(empty, **ds)
    assert int(mask.sum()) == 0 and bool((xyz == -1).all())
    big = torch.rand(64, 64, 1024, generator=torch.Generator().manual_seed(1)) * 2.4 - 1.2      # BASELINE batch
    xyz, mask = ops.backproject(big.cuda(), **ds)
    # size-independent property: |xyz| equals the decoded depth wherever the mask is set
    unit = (big.clamp(-1, 1) + 1) / 2
    depth = torch.exp2(unit * 5.84) - 1
    norm = xyz.norm(dim=1).cpu()
    m = mask.bool().cpu()
    assert torch.equal(m, (depth > 1.0) & (depth < 56.0))
    assert float(((norm - depth)[m]).abs().max()) < 1e-4


def test_attention_reference_value_margin(ops):
    """The bf16 softmax takes its running reference over every fourth column pair (attention.cu: sub-sampled maximum).  Any
    reference is exact as long as 2^(s - m) stays inside the fp32 / bf16 exponent range, i.e. while no score lies more than
    ~85 nats above the sampled maximum.  Plant outliers in UNSAMPLED key columns (2..7 of every 8) far above everything else -
    up to 60 nats, far outside anything a softmax over real logits produces - and compare with the fp32 reference."""
    B, heads, T, d = 1, 2, 512, 32
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B, heads * 3 * d, T, generator=g) * 0.5
    v = qkv.view(B, heads, 3, d, T)
    for head, (key, gap) in enumerate(((133, 30.0), (390, 60.0))):          # 133 % 8 = 5, 390 % 8 = 6: never sampled
        q_dir = torch.randn(d, generator=g)
        q_dir /= q_dir.norm()
        v[0, head, 0] = q_dir[:, None] * 4.0 + 0.05 * torch.randn(d, T, generator=g)      # every query points along q_dir
        v[0, head, 1, :, key] = q_dir * gap * d ** 0.5 / 4.0                                  # q . k / sqrt(d) ~ gap for that key
    y = ops.qkv_attention_legacy(qkv.cuda(), heads)
    ref = R.qkv_attention_legacy(qkv, heads)
    assert bool(torch.isfinite(y).all())
    assert rel(y, ref) < 1e-2


HALO_CASES = [
    # B, H, W, kh, kw, pad(l, r, t), residual scale (None: no residual)
    (2, 8, 256, 3, 3, (1, 1, 1), None),          # R2DM level-0 ring conv, small
    (2, 64, 1024, 3, 3, (1, 1, 1), 0.70710678),  # the same at 64 x 1024: ~7 tiles per CTA, every barrier phase several times
    (1, 4, 512, 1, 4, (1, 2, 0), 1.0),           # decoder curve-wise (1,4), asymmetric circular pad
    (1, 4, 128, 1, 5, (2, 2, 0), None),          # decoder Upsample conv (1,5): one tile per row
    (3, 16, 512, 3, 3, (1, 1, 1), 1.0),
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_halo_tile_conv_matches_streamed_kernel(ops, case):
    """64 -> 64 channel convolutions take the halo-tile kernel (gemm_halo.cu) in the models.  Run the conv as the models do
    (stored bf16 output + GroupNorm granule statistics) on both kernels: same K order per output element and same statistics
    order, so stored values and statistics must be the same bits; and the stored values sit within bf16 rounding of the oracle."""
    B, H, W, kh, kw, pad, rs = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = torch.randn(B, 64, H, W, generator=g)
    w = torch.randn(64, 64, kh, kw, generator=g) / (64 * kh * kw) ** 0.5
    b = torch.randn(64, generator=g)
    res = torch.randn(B, 64, H, W, generator=g) if rs is not None else None
    kw_ = dict(padding=pad, residual=res.cuda() if res is not None else None, res_scale=rs if rs is not None else 1.0)
    y1, g1 = ops.conv2d_stored(x.cuda(), w.cuda(), b.cuda(), halo_kernel=True, **kw_)
    y0, g0 = ops.conv2d_stored(x.cuda(), w.cuda(), b.cuda(), halo_kernel=False, **kw_)
    ref = R.circular_conv2d(x.bfloat16().float(), w.bfloat16().float(), b, (pad[0], pad[1], pad[2], kh - 1 - pad[2]))
    if res is not None:
        ref = ref + rs * res.bfloat16().float()
    assert rel(y0, ref) < 3e-3 and rel(y1, ref) < 3e-3
    assert torch.equal(y1, y0)
    assert torch.equal(g1, g0)
    # the statistics are those of the stored values: summed over the tiles they give the per-granule sums
    s = y1.reshape(B, 8, 8, H * W).sum(dim=(2, 3))
    assert torch.allclose(g1[..., 0].sum(dim=1), s, rtol=1e-3, atol=1e-1)


@pytest.mark.parametrize("sigma,tol", [(1.0, 1e-2), (3.0, 1e-2), (5.0, 5e-2)])
def test_attention_reference_regimes(ops, sigma, tol):
    """The bf16 softmax keeps reference 0 for rows whose first-tile maximum lies within 2^+-24 (raw scores go to ex2), otherwise
    a reference 2^40 above the tile maximum, and moves it, one tile late, when a later tile's maximum has left the range
    (attention.cu: streamed tiles, fast tiles).  sigma = 1: every row stays on reference 0; 3: logits of +-40 nats, rows start on
    either side and cross over; 5: +-100 nats, near one-hot rows whose maxima jump by tens of nats between tiles (there the bf16
    rounding of q and k alone moves a logit by +-0.3 nats, hence the wider bound).  All against the fp32 reference, all finite."""
    B, heads, T, d = 1, 4, 1024, 32
    g = torch.Generator().manual_seed(int(sigma * 10))
    qkv = torch.randn(B, heads * 3 * d, T, generator=g)
    v = qkv.view(B, heads, 3, d, T)
    v[:, :, 0] *= sigma
    v[:, :, 1] *= sigma
    y = ops.qkv_attention_legacy(qkv.cuda(), heads)
    ref = R.qkv_attention_legacy(qkv, heads)
    assert bool(torch.isfinite(y).all())
    assert rel(y, ref) < tol
