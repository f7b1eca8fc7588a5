"""TEST INFRASTRUCTURE ONLY (imported by tests/, bench.py's baseline legs and fixture generators).  CPU restatement of
the R2DM pixel-space denoiser of the reference (BASELINE config 5):
  EfficientUNet.forward                     lidm/modules/unets/efficient_unet.py:188-295 (Block :113-186,
                                            ResidualBlock :55-110, SelfAttentionBlock :23-52)
  ops.Conv2d / Pad / Resample / AdaGN /     lidm/modules/unets/ops.py:33-50, 52-143, 146-173, 176-200, 14-27
  SinusoidalPositionalEmbedding
  encoding.FourierFeatures / polar coords   lidm/modules/unets/encoding.py:93-105, 133-163
as plain functions over the module's state dict.  Pinned against the unmodified reference module by
tests/test_oracle_r2dm.py (fixtures tests/golden/r2dm_*.npz written by oracle/make_golden_r2dm.py).  The sampler around it
(R2DMDiffusion, lidm/models/diffusion/ddpm_r2dm.py:11-380) is the DDPM / DDIM eps-parameterisation arithmetic of
oracle/torch_ref.py with `timesteps: 1024` and no first stage."""
import math
from typing import Dict, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def ring_pad(h, left, right, top, bottom):
    """ops.Pad (ops.py:33-50) with ring=True: circular on W, zeros on H."""
    if left or right:
        h = F.pad(h, (left, right, 0, 0), mode="circular")
    if top or bottom:
        h = F.pad(h, (0, 0, top, bottom), mode="constant")
    return h


def conv2d_ring(sd, p, x, pad=1):
    """ops.Conv2d (ops.py:146-173): ring padding, then a padding-free conv."""
    if pad:
        x = ring_pad(x, pad, pad, pad, pad)
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"])


def resample(h, up=1, down=1, window=(1, 3, 3, 1)):
    """ops.Resample.forward (ops.py:52-143), direction 'hw', ring=True, normalize=True."""
    k = torch.tensor(window, dtype=torch.float32)
    k = k / k.sum()
    k = k * (up * up) ** (k.ndim / 2)
    n = len(window)
    if up > 1:
        p0, p1 = (n - up + 1) // 2 + up - 1, (n - up) // 2
    else:
        p0, p1 = (n - down + 1) // 2, (n - down) // 2
    margin = max(p0, p1)
    h = F.pad(h, (margin, margin, 0, 0), mode="circular")
    h = F.pad(h, (0, 0, margin, margin), mode="constant")
    B, C, H, W = h.shape
    h = h.view(B, C, H, 1, W, 1)
    h = F.pad(h, [0, up - 1, 0, 0, 0, up - 1])
    h = h.view(B, C, H * up, W * up)
    h = h[..., margin * up - p0:(H - margin) * up + p1, margin * up - p0:(W - margin) * up + p1]
    kern = k[None, None].repeat(C, 1, 1).to(h.dtype)
    h = F.conv2d(h, kern[..., None, :], groups=C)
    h = F.conv2d(h, kern[..., :, None], groups=C)
    return h[:, :, ::down, ::down]


def sinusoidal_embedding(t, channels, max_period=10_000):
    """ops.SinusoidalPositionalEmbedding (ops.py:14-27): [sin | cos], frequencies exp(-ln(P) i / (half - 1))."""
    h = -np.log(max_period) / (channels // 2 - 1)
    h = torch.exp(h * torch.arange(channels // 2, device=t.device))
    h = t[:, None] * h[None, :]
    return torch.cat([h.sin(), h.cos()], dim=-1).to(t)


def polar_coords(H, W):
    """encoding.generate_polar_coords (encoding.py:93-105): elevation 10..-30 deg, azimuth 180..-180 deg, radians."""
    elevation = (1 - torch.arange(H) / H) * (10 - (-30)) + (-30)
    azimuth = (1 - torch.arange(W) / W) * (180 - (-180)) + (-180)
    elevation, azimuth = torch.meshgrid([elevation, azimuth], indexing="ij")
    return torch.stack([elevation, azimuth])[None].deg2rad()


def fourier_features(H, W):
    """encoding.FourierFeatures (encoding.py:133-163) applied to the polar coordinates: (1, 2 (L_h + L_w), H, W)."""
    L_h, L_w = int(np.ceil(np.log2(H))), int(np.ceil(np.log2(W)))
    fh = torch.cat([torch.arange(L_h).exp2(), torch.zeros(L_w)])
    fw = torch.cat([torch.zeros(L_h), torch.arange(L_w).exp2()])
    freqs = torch.stack([fh, fw], dim=-1)[..., None, None]
    c = F.conv2d(polar_coords(H, W), weight=freqs, bias=torch.zeros(len(fh)))
    return torch.cat([c.sin(), c.cos()], dim=1)


def _gn(x, groups, eps, w=None, b=None):
    return F.group_norm(x, groups, w, b, eps)


def residual_block(sd, p, x, temb, groups, eps, scale=1 / np.sqrt(2)):
    """ResidualBlock.forward (efficient_unet.py:55-110): GN -> SiLU -> conv; AdaGN(temb) -> SiLU -> conv; (skip + h) * scale."""
    h = F.silu(_gn(x, groups, eps, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"]))
    h = conv2d_ring(sd, p + ".conv1", h)
    e = F.linear(F.silu(temb), sd[p + ".norm2.proj.1.weight"], sd[p + ".norm2.proj.1.bias"])[:, :, None, None]
    sc, sh = e.chunk(2, dim=1)
    h = _gn(h, groups, eps) * (1 + sc) + sh                      # AdaGN (ops.py:176-200): GroupNorm without affine
    h = conv2d_ring(sd, p + ".conv2", F.silu(h))
    sk = conv2d_ring(sd, p + ".skip", x, pad=0) if (p + ".skip.weight") in sd else x
    return (sk + h) * torch.tensor(scale).float()


def self_attention_block(sd, p, x, heads, groups, eps, scale=1 / np.sqrt(2)):
    """SelfAttentionBlock.forward (efficient_unet.py:23-52): GN -> nn.MultiheadAttention(batch_first) -> (x + h) * scale."""
    B, C, H, W = x.shape
    h = _gn(x, groups, eps, sd[p + ".norm.weight"], sd[p + ".norm.bias"]).reshape(B, C, H * W).permute(0, 2, 1)
    qkv = F.linear(h, sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"])
    q, k, v = (t.reshape(B, H * W, heads, C // heads).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(C // heads), dim=-1) @ v
    a = a.permute(0, 2, 1, 3).reshape(B, H * W, C)
    a = F.linear(a, sd[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"])
    return (x + a.permute(0, 2, 1).reshape(B, C, H, W)) * torch.tensor(scale).float()


def block(sd, p, h, temb, n_res, groups, eps, heads, attn=False, up=1, down=1):
    """Block.forward (efficient_unet.py:113-186): [conv + FIR down] -> residual blocks -> [attention] -> [FIR up + conv]."""
    if down > 1:
        h = resample(conv2d_ring(sd, p + ".downsample.0", h), down=down)
    for i in range(n_res):
        h = residual_block(sd, f"{p}.residual_blocks.{i}", h, temb, groups, eps)
    if attn:
        h = self_attention_block(sd, p + ".self_attn_block", h, heads, groups, eps)
    if up > 1:
        h = conv2d_ring(sd, p + ".upsample.1", resample(h, up=up))
    return h


@torch.no_grad()
def efficient_unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, timesteps: torch.Tensor, *, resolution: Sequence[int],
                           base_channels: int = 64, channel_multiplier: Sequence[int] = (1, 2, 4, 8),
                           num_residual_blocks: Sequence[int] = (3, 3, 3, 3), gn_num_groups: int = 8, gn_eps: float = 1e-6,
                           attn_num_heads: int = 8) -> torch.Tensor:
    """EfficientUNet.forward (efficient_unet.py:262-295), coords_encoding 'fourier_features', ring=True."""
    temb = sinusoidal_embedding(timesteps.to(x), base_channels)
    temb = F.linear(temb, sd["time_embedding.1.weight"], sd["time_embedding.1.bias"])
    temb = F.linear(F.silu(temb), sd["time_embedding.3.weight"], sd["time_embedding.3.bias"])
    cenc = fourier_features(*resolution).repeat_interleave(x.shape[0], dim=0)
    h = conv2d_ring(sd, "in_conv", torch.cat([x, cenc], dim=1))
    N = list(num_residual_blocks)
    kw = dict(groups=gn_num_groups, eps=gn_eps, heads=attn_num_heads)
    h1 = block(sd, "d_block1", h, temb, N[0], **kw)
    h2 = block(sd, "d_block2", h1, temb, N[1], down=2, **kw)
    h3 = block(sd, "d_block3", h2, temb, N[2], down=2, **kw)
    h4 = block(sd, "d_block4", h3, temb, N[3], down=2, attn=True, **kw)
    h = block(sd, "u_block4", h4, temb, N[3], up=2, attn=True, **kw)
    h = block(sd, "u_block3", torch.cat([h, h3], dim=1), temb, N[2], up=2, **kw)
    h = block(sd, "u_block2", torch.cat([h, h2], dim=1), temb, N[1], up=2, **kw)
    h = block(sd, "u_block1", torch.cat([h, h1], dim=1), temb, N[0], **kw)
    return conv2d_ring(sd, "out_conv", h)
