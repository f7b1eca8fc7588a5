"""Operator-level entry points (same kernels the model uses), torch CUDA tensors in / out.
Each mirrors one reference primitive; see include/lidm_b200.h for the file:line citations."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib
from .engine import _f32c, _stream_ptr


def circular_conv2d(x, weight, bias=None, padding: Optional[Sequence[int]] = None, stride: int = 1, residual=None):
    """CircularConv2d.forward (reference lidm/modules/basic.py:52-59).  padding = (left, right, top, bottom)."""
    x, weight = _f32c(x, "x"), _f32c(weight, "weight")
    B, Cin, H, W = x.shape
    Cout, Cin2, kh, kw = weight.shape
    assert Cin == Cin2
    pl, pr, pt, pb = padding if padding is not None else (0, 0, 0, 0)
    Ho = (H + pt + pb - kh) // stride + 1
    Wo = (W + pl + pr - kw) // stride + 1
    out = torch.empty((B, Cout, Ho, Wo), dtype=torch.float32, device=x.device)
    b = _f32c(bias, "bias") if bias is not None else None
    r = _f32c(residual, "residual") if residual is not None else None
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.lidm_op_circular_conv2d(x.data_ptr(), B, Cin, H, W, weight.data_ptr(),
                                               b.data_ptr() if b is not None else None, Cout, kh, kw, pl, pr, pt, pb,
                                               stride, r.data_ptr() if r is not None else None, out.data_ptr(),
                                               _stream_ptr(x.device)))
    return out


def conv2d_stored(x, weight, bias=None, padding=(1, 1, 1), residual=None, res_scale=1.0, halo_kernel=False):
    """The convolution as the model runs it: bf16 channels-last stored output plus the GroupNorm granule statistics the epilogue
    writes.  padding = (left, right, top); returns (out fp32 NCHW, gst (B, H*W/128, Cout/8, 2)).  halo_kernel picks the halo-tile
    kernel (csrc/gemm_halo.cu) instead of the streamed implicit GEMM; both must produce the same bits."""
    x, weight = _f32c(x, "x"), _f32c(weight, "weight")
    B, Cin, H, W = x.shape
    Cout, _, kh, kw = weight.shape
    pl, pr, pt = padding
    out = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    gst = torch.empty((B, H * W // 128, Cout // 8, 2), dtype=torch.float32, device=x.device)
    b = _f32c(bias, "bias") if bias is not None else None
    r = _f32c(residual, "residual") if residual is not None else None
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.lidm_op_conv2d_stored(x.data_ptr(), B, Cin, H, W, weight.data_ptr(), b.data_ptr() if b is not None else None,
                                             Cout, kh, kw, pl, pr, pt, r.data_ptr() if r is not None else None, float(res_scale),
                                             int(halo_kernel), out.data_ptr(), gst.data_ptr(), _stream_ptr(x.device)))
    return out, gst


def group_norm(x, gamma, beta, eps=1e-5, groups=32, silu=False):
    """GroupNorm32 (+ SiLU) (reference lidm/modules/basic.py:339-341)."""
    x, gamma, beta = _f32c(x, "x"), _f32c(gamma, "gamma"), _f32c(beta, "beta")
    B, C, H, W = x.shape
    out = torch.empty_like(x)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.lidm_op_groupnorm(x.data_ptr(), B, C, H, W, gamma.data_ptr(), beta.data_ptr(), float(eps),
                                         groups, int(silu), out.data_ptr(), _stream_ptr(x.device)))
    return out


def qkv_attention_legacy(qkv, n_heads):
    """QKVAttentionLegacy.forward (reference lidm/modules/diffusion/openaimodel.py:358-374), head dim 32."""
    qkv = _f32c(qkv, "qkv")
    B, width, T = qkv.shape
    assert width == n_heads * 96, "head dim must be 32"
    out = torch.empty((B, n_heads * 32, T), dtype=torch.float32, device=qkv.device)
    lib = _lib.load()
    with torch.cuda.device(qkv.device):
        _lib.check(lib.lidm_op_qkv_attention_legacy(qkv.data_ptr(), B, n_heads, T, out.data_ptr(),
                                                    _stream_ptr(qkv.device)))
    return out


def ddim_step(x, e_t, coef, noise=None, temperature=1.0):
    """p_sample_ddim update (reference lidm/models/diffusion/ddim.py:191-206).  coef = (a_t, a_prev, sigma_t,
    sqrt_one_minus_at).  Returns (x_prev, pred_x0)."""
    x, e_t = _f32c(x, "x"), _f32c(e_t, "e_t")
    nz = _f32c(noise, "noise") if noise is not None else None
    x_prev, pred_x0 = torch.empty_like(x), torch.empty_like(x)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.lidm_ddim_step(x.data_ptr(), e_t.data_ptr(), nz.data_ptr() if nz is not None else None,
                                      float(coef[0]), float(coef[1]), float(coef[2]), float(coef[3]),
                                      float(temperature), x_prev.data_ptr(), pred_x0.data_ptr(), x.numel(),
                                      _stream_ptr(x.device)))
    return x_prev, pred_x0


def ddpm_step(x, eps, noise, coef, clip_denoised=False, return_x0=False):
    """Ancestral DDPM update (reference ddpm.py:1090-1119 after the U-Net) in one kernel.  coef: (B, 5) fp32 CUDA tensor of
    per-sample (sqrt_recip_ac, sqrt_recipm1_ac, posterior_mean_coef1, posterior_mean_coef2, (t != 0) * exp(0.5 logvar)).
    Returns x_prev (and x_recon).  Bit-identical to the eager tensor expression."""
    x, eps, noise = _f32c(x, "x"), _f32c(eps, "eps"), _f32c(noise, "noise")
    coef = _f32c(coef, "coef")
    B = x.shape[0]
    if coef.shape != (B, 5) or eps.shape != x.shape or noise.shape != x.shape:
        raise ValueError("ddpm_step: eps / noise must match x and coef must be (B, 5)")
    x_prev = torch.empty_like(x)
    x0 = torch.empty_like(x) if return_x0 else None
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.lidm_ddpm_step(x.data_ptr(), eps.data_ptr(), noise.data_ptr(), coef.data_ptr(), B, x.numel() // B,
                                      int(bool(clip_denoised)), x_prev.data_ptr(), x0.data_ptr() if x0 is not None else None,
                                      _stream_ptr(x.device)))
    return (x_prev, x0) if return_x0 else x_prev


def backproject(img, fov, depth_range, depth_scale, log_scale=True, return_mask=True, input_is_unit=False):
    """range2xyz on the device: img (B,H,W) or (B,1,H,W) fp32 in [-1,1] -> xyz (B,3,H,W) fp32 (-1 where masked),
    mask (B,H,W) uint8 (reference lidm/utils/lidar_utils.py:175-204 + scripts/sample.py:29-35)."""
    img = _f32c(img, "img")
    if img.dim() == 4:
        assert img.shape[1] == 1
        img = img[:, 0]
    img = img.contiguous()
    B, H, W = img.shape
    xyz = torch.empty((B, 3, H, W), dtype=torch.float32, device=img.device)
    mask = torch.empty((B, H, W), dtype=torch.uint8, device=img.device) if return_mask else None
    lib = _lib.load()
    with torch.cuda.device(img.device):
        _lib.check(lib.lidm_backproject(img.data_ptr(), B, H, W, float(fov[0]), float(fov[1]), float(depth_range[0]),
                                        float(depth_range[1]), float(depth_scale), int(bool(log_scale)),
                                        int(bool(input_is_unit)), xyz.data_ptr(), mask.data_ptr() if mask is not None else None,
                                        _stream_ptr(img.device)))
    return (xyz, mask) if return_mask else xyz


def to_uint8_image(x):
    """custom_to_pil's array (reference scripts/sample.py:38-45): uint8(255 * (clip(x,-1,1)+1)/2), same shape as x."""
    x = _f32c(x, "x")
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.lidm_to_uint8_image(x.data_ptr(), out.data_ptr(), x.numel(), _stream_ptr(x.device)))
    return out


def compact_points(xyz, mask):
    """`pcd[mask, :]` for a batch (reference lidm/utils/lidar_utils.py:169-171): xyz (B,3,H,W) fp32, mask (B,H,W) uint8 ->
    (points (B, H*W, 3) fp32 with each sample's valid points packed first in pixel order, counts (B,) int32)."""
    xyz = _f32c(xyz, "xyz")
    if mask.dtype != torch.uint8 or not mask.is_cuda:
        raise ValueError("mask must be a CUDA uint8 tensor")
    mask = mask.contiguous()
    B, _, H, W = xyz.shape
    points = torch.empty((B, H * W, 3), dtype=torch.float32, device=xyz.device)
    counts = torch.empty((B,), dtype=torch.int32, device=xyz.device)
    lib = _lib.load()
    with torch.cuda.device(xyz.device):
        _lib.check(lib.lidm_compact_points(xyz.data_ptr(), mask.data_ptr(), B, H * W, points.data_ptr(), counts.data_ptr(),
                                           _stream_ptr(xyz.device)))
    return points, counts
