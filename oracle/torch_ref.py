"""ORACLE (test infrastructure, not product): CPU restatement of the reference's sampling path.

Plain torch.nn.functional / numpy fp32 (fp64 where the reference uses it), one function per reference
function, each citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs may import this module; the product
(lidar_layout_b200/) never does and fails loudly without its CUDA extension.

Pinning: the reference ships no golden vectors or tests for this path, so the
oracle is pinned against the *reference itself executed in the build container* through
oracle/ref_shim.py: oracle/make_golden.py loads the synthetic state-dict into the real reference
modules, runs them, and commits small fixtures under tests/golden/; tests/test_oracle_golden.py checks
this file against those fixtures on any machine, and tests/test_oracle_vs_reference.py compares
directly when /root/reference is present.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from lidar_layout_b200.config import AEConfig, LidmConfig, UNetConfig
from lidar_layout_b200.weights import (AE_PREFIX, DOWNSAMPLE_STRIDE2PAD, UNET_PREFIX, UPSAMPLE_STRIDE2KERNEL,
                                       decoder_levels, encoder_levels, unet_blocks)

# --------------------------------------------------------------------------------------
# schedules
# --------------------------------------------------------------------------------------


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2):
    """lidm/modules/basic.py:147-169 (linear branch; float64)."""
    if schedule != "linear":
        raise ValueError(schedule)
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    return betas.numpy()


def register_schedule(cfg: LidmConfig) -> Dict[str, torch.Tensor]:
    """DDPM.register_schedule, lidm/models/diffusion/ddpm.py:120-160 (float64 numpy -> float32 buffers)."""
    betas = make_beta_schedule(cfg.beta_schedule, cfg.timesteps, cfg.linear_start, cfg.linear_end)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    t = lambda a: torch.tensor(a, dtype=torch.float32)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return dict(
        betas=t(betas), alphas_cumprod=t(ac), alphas_cumprod_prev=t(ac_prev),
        sqrt_alphas_cumprod=t(np.sqrt(ac)), sqrt_one_minus_alphas_cumprod=t(np.sqrt(1.0 - ac)),
        sqrt_recip_alphas_cumprod=t(np.sqrt(1.0 / ac)), sqrt_recipm1_alphas_cumprod=t(np.sqrt(1.0 / ac - 1)),
        posterior_variance=t(post_var),
        posterior_log_variance_clipped=t(np.log(np.maximum(post_var, 1e-20))),
        posterior_mean_coef1=t(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        posterior_mean_coef2=t((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
    )


def make_ddim_timesteps(num_ddim_timesteps, num_ddpm_timesteps):
    """lidm/modules/basic.py:172-185 ('uniform')."""
    c = num_ddpm_timesteps // num_ddim_timesteps
    return np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1


def make_ddim_sampling_parameters(alphacums: torch.Tensor, ddim_timesteps, eta):
    """lidm/modules/basic.py:188-197.  alphacums is the float32 CPU tensor; note the mixed dtypes the
    reference produces: alphas = torch f32 tensor, alphas_prev = numpy f64 (python floats of f32 values),
    sigmas = eta * np.sqrt(...) evaluated on a torch tensor -> torch f64."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def ddim_schedule(cfg: LidmConfig, S: int, eta: float):
    """DDIMSampler.make_schedule (lidm/models/diffusion/ddim.py:26-55) + the f32 rounding done by
    torch.full in p_sample_ddim (:191-194).  Returns timesteps (int64 ndarray, ascending) and an (n,4)
    float32 ndarray [a_t, a_prev, sigma_t, sqrt_one_minus_at] indexed by `index`."""
    sched = register_schedule(cfg)
    ts = make_ddim_timesteps(S, cfg.timesteps)
    sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(sched["alphas_cumprod"], ts, eta)
    sqrt_1m = np.sqrt(1.0 - alphas)  # np.sqrt on a torch f32 tensor -> torch f32
    n = len(ts)
    table = np.zeros((n, 4), dtype=np.float32)
    for i in range(n):
        table[i, 0] = torch.full((1,), alphas[i]).item()
        table[i, 1] = torch.full((1,), float(alphas_prev[i])).item()
        table[i, 2] = torch.full((1,), float(sigmas[i])).item()
        table[i, 3] = torch.full((1,), sqrt_1m[i]).item()
    return ts.astype(np.int64), table


# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------


def circular_conv2d(x, w, b, pad=None, stride=1):
    """CircularConv2d.forward, lidm/modules/basic.py:52-59: circular pad on W, zero pad on H, conv pad 0.
    pad = (left, right, top, bottom) or None (no padding)."""
    if pad is not None:
        h1, h2, v1, v2 = pad
        if h1 + h2 > 0:
            x = F.pad(x, (h1, h2, 0, 0), mode="circular")
        if v1 + v2 > 0:
            x = F.pad(x, (0, 0, v1, v2), mode="constant")
    return F.conv2d(x, w, b, stride=stride)


def timestep_embedding(timesteps, dim, max_period=10000):
    """lidm/modules/basic.py:278-296 (cos first, then sin)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def group_norm(x, w, b, eps, groups=32):
    return F.group_norm(x.float(), groups, w, b, eps).type(x.dtype)


def silu(x):
    return x * torch.sigmoid(x)


# --------------------------------------------------------------------------------------
# U-Net (openaimodel.py)
# --------------------------------------------------------------------------------------


def _resblock(sd, p, x, emb):
    """ResBlock._forward, lidm/modules/diffusion/openaimodel.py:256-276 (no updown, no scale-shift)."""
    h = group_norm(x, sd[p + ".in_layers.0.weight"], sd[p + ".in_layers.0.bias"], 1e-5)
    h = F.silu(h)
    h = circular_conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], (1, 1, 1, 1))
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = group_norm(h, sd[p + ".out_layers.0.weight"], sd[p + ".out_layers.0.bias"], 1e-5)
    h = F.silu(h)
    h = circular_conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], (1, 1, 1, 1))
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def qkv_attention_legacy(qkv, n_heads):
    """QKVAttentionLegacy.forward, openaimodel.py:358-374."""
    bs, width, length = qkv.shape
    ch = width // (3 * n_heads)
    q, k, v = qkv.reshape(bs * n_heads, ch * 3, length).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    weight = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    weight = torch.softmax(weight.float(), dim=-1).type(weight.dtype)
    a = torch.einsum("bts,bcs->bct", weight, v)
    return a.reshape(bs, -1, length)


def _attnblock(sd, p, x, heads):
    """AttentionBlock._forward, openaimodel.py:320-326."""
    b, c, *spatial = x.shape
    x = x.reshape(b, c, -1)
    qkv = F.conv1d(group_norm(x, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5),
                   sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    h = qkv_attention_legacy(qkv, heads)
    h = F.conv1d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (x + h).reshape(b, c, *spatial)


def layer_norm(x, w, b, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def _cross_attention(sd, p, x, context, heads):
    """CrossAttention.forward, lidm/modules/attention.py:170-193 (mask None; to_q/k/v have no bias;
    scale = dim_head ** -0.5, :158; softmax in the input dtype)."""
    q = F.linear(x, sd[p + ".to_q.weight"])
    context = x if context is None else context
    k = F.linear(context, sd[p + ".to_k.weight"])
    v = F.linear(context, sd[p + ".to_v.weight"])
    b, n, inner = q.shape
    d = inner // heads

    def split(t):   # 'b n (h d) -> (b h) n d'
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", attn, v)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, inner)   # '(b h) n d -> b n (h d)'
    return F.linear(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def _basic_transformer_block(sd, p, x, context, heads):
    """BasicTransformerBlock._forward, lidm/modules/attention.py:211-215; FeedForward with GEGLU (:36-65)."""
    x = _cross_attention(sd, p + ".attn1", layer_norm(x, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"]), None, heads) + x
    x = _cross_attention(sd, p + ".attn2", layer_norm(x, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"]), context, heads) + x
    h = layer_norm(x, sd[p + ".norm3.weight"], sd[p + ".norm3.bias"])
    h = F.linear(h, sd[p + ".ff.net.0.proj.weight"], sd[p + ".ff.net.0.proj.bias"])
    a, gate = h.chunk(2, dim=-1)
    h = a * F.gelu(gate)
    h = F.linear(h, sd[p + ".ff.net.2.weight"], sd[p + ".ff.net.2.bias"])
    return h + x


def _spatial_transformer(sd, p, x, context, heads, depth):
    """SpatialTransformer.forward, lidm/modules/attention.py:250-261 (Normalize = GroupNorm(32, eps 1e-6), :76)."""
    b, c, h, w = x.shape
    x_in = x
    x = group_norm(x, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    x = F.conv2d(x, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    x = x.reshape(b, c, h * w).permute(0, 2, 1)
    for d in range(depth):
        x = _basic_transformer_block(sd, f"{p}.transformer_blocks.{d}", x, context, heads)
    x = x.permute(0, 2, 1).reshape(b, c, h, w)
    x = F.conv2d(x, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x + x_in


def _run_block(sd, prefix, layers, h, emb, context=None):
    for j, layer in enumerate(layers):
        p = f"{prefix}.{j}"
        kind = layer[0]
        if kind == "conv":
            h = circular_conv2d(h, sd[p + ".weight"], sd[p + ".bias"], (1, 1, 1, 1))
        elif kind == "res":
            h = _resblock(sd, p, h, emb)
        elif kind == "attn":
            h = _attnblock(sd, p, h, layer[2])
        elif kind == "st":
            h = _spatial_transformer(sd, p, h, context, layer[2], layer[3])
        elif kind == "down":   # Downsample.forward openaimodel.py:159-161
            h = circular_conv2d(h, sd[p + ".op.weight"], sd[p + ".op.bias"], (1, 1, 1, 1), stride=2)
        elif kind == "up":     # Upsample.forward openaimodel.py:108-118
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = circular_conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], (1, 1, 1, 1))
    return h


@torch.no_grad()
def unet_forward(sd: Dict[str, torch.Tensor], cfg: UNetConfig, x: torch.Tensor, timesteps: torch.Tensor,
                 context: Optional[torch.Tensor] = None, prefix: str = UNET_PREFIX) -> torch.Tensor:
    """UNetModel.forward, lidm/modules/diffusion/openaimodel.py:719-751 (context = cross-attention conditioning)."""
    inputs, middle, outputs, _ = unet_blocks(cfg)
    t_emb = timestep_embedding(timesteps, cfg.model_channels)
    emb = F.linear(t_emb, sd[prefix + "time_embed.0.weight"], sd[prefix + "time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd[prefix + "time_embed.2.weight"], sd[prefix + "time_embed.2.bias"])
    hs = []
    h = x.float()
    for i, layers in enumerate(inputs):
        h = _run_block(sd, f"{prefix}input_blocks.{i}", layers, h, emb, context)
        hs.append(h)
    h = _run_block(sd, f"{prefix}middle_block", middle, h, emb, context)
    for i, layers in enumerate(outputs):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sd, f"{prefix}output_blocks.{i}", layers, h, emb, context)
    h = F.silu(group_norm(h, sd[prefix + "out.0.weight"], sd[prefix + "out.0.bias"], 1e-5))
    return circular_conv2d(h, sd[prefix + "out.2.weight"], sd[prefix + "out.2.bias"], (1, 1, 1, 1))


# --------------------------------------------------------------------------------------
# first stage (autoencoder.py, model_lidm.py, taming quantiser as vq.py)
# --------------------------------------------------------------------------------------

UNIFORM_KERNEL2PAD = {(3, 3): (1, 1, 1, 1), (1, 4): (1, 2, 0, 0)}
UPSAMPLE_STRIDE2PAD = {(1, 2): (2, 2, 0, 0), (1, 4): (3, 3, 0, 0), (2, 1): (0, 0, 2, 2), (2, 2): (1, 1, 1, 1)}


def vq_quantize(z: torch.Tensor, codebook: torch.Tensor):
    """taming VectorQuantizer2.forward (eval arithmetic), as lidm/models/ae/vq.py:66-79 with the
    b c h w -> b h w c permute.  Returns (z_q (B,C,H,W), indices (B*H*W,) int64)."""
    zp = z.permute(0, 2, 3, 1).contiguous()
    zf = zp.view(-1, codebook.shape[1])
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(codebook ** 2, dim=1) \
        - 2 * torch.einsum("bd,dn->bn", zf, codebook.t())
    idx = torch.argmin(d, dim=1)
    zq = F.embedding(idx, codebook).view(zp.shape)
    zq = zp + (zq - zp)          # straight-through form kept: it is the value the reference decodes
    return zq.permute(0, 3, 1, 2).contiguous(), idx


def _resnet_block(sd, p, x, kernel):
    """ResnetBlock.forward, lidm/modules/diffusion/model_lidm.py:127-147 (temb is None)."""
    pad = UNIFORM_KERNEL2PAD[kernel]
    h = group_norm(x, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-6)
    h = silu(h)
    h = circular_conv2d(h, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], pad)
    h = group_norm(h, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-6)
    h = silu(h)
    h = circular_conv2d(h, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], pad)
    if (p + ".nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def _attn_block(sd, p, x):
    """AttnBlock.forward, model_lidm.py:184-208 (single head, scale C^-1/2)."""
    h_ = group_norm(x, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    q = F.conv2d(h_, sd[p + ".q.weight"], sd[p + ".q.bias"])
    k = F.conv2d(h_, sd[p + ".k.weight"], sd[p + ".k.bias"])
    v = F.conv2d(h_, sd[p + ".v.weight"], sd[p + ".v.bias"])
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    h_ = F.conv2d(h_, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x + h_


@torch.no_grad()
def decoder_forward(sd, cfg: AEConfig, z, prefix: str = AE_PREFIX + "decoder."):
    """Decoder.forward, model_lidm.py:385-417."""
    top, levels, _ = decoder_levels(cfg)
    h = circular_conv2d(z, sd[prefix + "conv_in.weight"], sd[prefix + "conv_in.bias"], (1, 1, 1, 1))
    h = _resnet_block(sd, prefix + "mid.block_1", h, (3, 3))
    h = _attn_block(sd, prefix + "mid.attn_1", h)
    h = _resnet_block(sd, prefix + "mid.block_2", h, (3, 3))
    for i_level in reversed(range(len(cfg.ch_mult))):
        lv = levels[i_level]
        for i_block in range(cfg.num_res_blocks + 1):
            h = _resnet_block(sd, prefix + f"up.{i_level}.block.{i_block}", h, lv["kernel"])
        if i_level != 0:
            stride = lv["stride"]
            h = F.interpolate(h, scale_factor=stride, mode="bilinear", align_corners=True)   # :57-61
            p = prefix + f"up.{i_level}.upsample.conv"
            h = circular_conv2d(h, sd[p + ".weight"], sd[p + ".bias"], UPSAMPLE_STRIDE2PAD[stride])
    h = silu(group_norm(h, sd[prefix + "norm_out.weight"], sd[prefix + "norm_out.bias"], 1e-6))
    return circular_conv2d(h, sd[prefix + "conv_out.weight"], sd[prefix + "conv_out.bias"], (1, 2, 0, 0))


@torch.no_grad()
def encoder_forward(sd, cfg: AEConfig, x, prefix: str = AE_PREFIX + "encoder."):
    """Encoder.forward, model_lidm.py:284-312 (ResnetBlocks with the default 3x3 kernel; Downsample = strided
    CircularConv2d with the asymmetric pads of DOWNSAMPLE_STRIDE2PAD_DICT, :64-65)."""
    levels, _ = encoder_levels(cfg)
    h = circular_conv2d(x, sd[prefix + "conv_in.weight"], sd[prefix + "conv_in.bias"], (1, 1, 1, 1))
    for i_level, lv in enumerate(levels):
        for i_block in range(cfg.num_res_blocks):
            h = _resnet_block(sd, prefix + f"down.{i_level}.block.{i_block}", h, (3, 3))
        if lv["stride"] is not None:
            p = prefix + f"down.{i_level}.downsample.conv"
            h = circular_conv2d(h, sd[p + ".weight"], sd[p + ".bias"], DOWNSAMPLE_STRIDE2PAD[lv["stride"]], stride=lv["stride"])
    h = _resnet_block(sd, prefix + "mid.block_1", h, (3, 3))
    h = _attn_block(sd, prefix + "mid.attn_1", h)
    h = _resnet_block(sd, prefix + "mid.block_2", h, (3, 3))
    h = silu(group_norm(h, sd[prefix + "norm_out.weight"], sd[prefix + "norm_out.bias"], 1e-6))
    return circular_conv2d(h, sd[prefix + "conv_out.weight"], sd[prefix + "conv_out.bias"], (1, 1, 1, 1))


@torch.no_grad()
def encode_first_stage(sd, cfg: LidmConfig, x):
    """LatentDiffusion.encode_first_stage (ddpm.py:837-...) -> VQModelInterface.encode (autoencoder.py:285-288):
    quant_conv(encoder(x)), NOT quantised (the quantiser sits in decode); scale via get_first_stage_encoding."""
    h = encoder_forward(sd, cfg.ae, x)
    return F.conv2d(h, sd[AE_PREFIX + "quant_conv.weight"], sd[AE_PREFIX + "quant_conv.bias"])


def get_first_stage_encoding(cfg: LidmConfig, z):
    """ddpm.py:546-556 for a tensor posterior: scale_factor * z."""
    return cfg.scale_factor * z


@torch.no_grad()
def decode_first_stage(sd, cfg: LidmConfig, z, force_not_quantize=False, return_indices=False):
    """LatentDiffusion.decode_first_stage (ddpm.py:717-775) -> VQModelInterface.decode
    (lidm/models/ae/autoencoder.py:290-302)."""
    z = 1.0 / cfg.scale_factor * z
    idx = None
    if not force_not_quantize:
        z, idx = vq_quantize(z, sd[AE_PREFIX + "quantize.embedding.weight"])
    q = F.conv2d(z, sd[AE_PREFIX + "post_quant_conv.weight"], sd[AE_PREFIX + "post_quant_conv.bias"])
    dec = decoder_forward(sd, cfg.ae, q)
    if cfg.ae.use_mask:
        mask = dec[:, 1:2] < 0.0
        dec = dec[:, 0:1]
        dec[mask] = -1.0
    return (dec, idx) if return_indices else dec


# --------------------------------------------------------------------------------------
# DDIM (ddim.py)
# --------------------------------------------------------------------------------------


def ddim_step(x, e_t, coef, noise=None, temperature=1.0, quantize=None):
    """DDIMSampler.p_sample_ddim arithmetic, lidm/models/diffusion/ddim.py:191-206.
    coef = (a_t, a_prev, sigma_t, sqrt_one_minus_at) float32 scalars.  quantize: optional callable applied to pred_x0
    (quantize_denoised, ddim.py:198-199)."""
    b = x.shape[0]
    a_t = torch.full((b, 1, 1, 1), float(coef[0]), device=x.device)
    a_prev = torch.full((b, 1, 1, 1), float(coef[1]), device=x.device)
    sigma_t = torch.full((b, 1, 1, 1), float(coef[2]), device=x.device)
    sqrt_one_minus_at = torch.full((b, 1, 1, 1), float(coef[3]), device=x.device)
    pred_x0 = (x - sqrt_one_minus_at * e_t) / a_t.sqrt()
    if quantize is not None:
        pred_x0 = quantize(pred_x0)
    dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
    n = torch.zeros_like(x) if noise is None else noise
    x_prev = a_prev.sqrt() * pred_x0 + dir_xt + sigma_t * n * temperature
    return x_prev, pred_x0


@torch.no_grad()
def apply_model(sd, cfg: LidmConfig, x, t, cond=None):
    """LatentDiffusion.apply_model (ddpm.py:900-1000) -> DiffusionWrapper.forward (ddpm.py:2313-2339) for the
    conditioning keys None / 'concat' / 'crossattn' with a single conditioning tensor."""
    if cfg.conditioning_key is None or cond is None:
        return unet_forward(sd, cfg.unet, x, t)
    if cfg.conditioning_key == "concat":
        return unet_forward(sd, cfg.unet, torch.cat([x, cond], dim=1), t)
    if cfg.conditioning_key == "crossattn":
        return unet_forward(sd, cfg.unet, x, t, context=cond)
    raise NotImplementedError(cfg.conditioning_key)


@torch.no_grad()
def guided_eps(sd, cfg: LidmConfig, x, t, cond, unconditional_conditioning=None, unconditional_guidance_scale=1.0):
    """The eps prediction of DDIMSampler.p_sample_ddim incl. classifier-free guidance, ddim.py:173-180."""
    if unconditional_conditioning is None or unconditional_guidance_scale == 1.0:
        return apply_model(sd, cfg, x, t, cond)
    x_in = torch.cat([x] * 2)
    t_in = torch.cat([t] * 2)
    c_in = torch.cat([unconditional_conditioning, cond])
    e_u, e_c = apply_model(sd, cfg, x_in, t_in, c_in).chunk(2)
    return e_u + unconditional_guidance_scale * (e_c - e_u)


@torch.no_grad()
def q_sample(cfg: LidmConfig, x_start, t, noise):
    """DDPM.q_sample, ddpm.py:306-309."""
    sched = register_schedule(cfg)
    return (_extract(sched["sqrt_alphas_cumprod"], t, x_start.shape) * x_start +
            _extract(sched["sqrt_one_minus_alphas_cumprod"], t, x_start.shape) * noise)


@torch.no_grad()
def ddim_sample(sd, cfg: LidmConfig, S, x_T, eta=0.0, noise=None, temperature=1.0, record=None, cond=None,
                unconditional_conditioning=None, unconditional_guidance_scale=1.0, quantize_x0=False, mask=None, x0=None,
                q_noise=None):
    """DDIMSampler.sample / ddim_sampling, ddim.py:57-165.
    noise: optional (n_steps,B,C,H,W) pre-generated tensor used for the sigma_t * randn term (in loop order).
    record: optional list; receives (x_t, t, eps, pred_x0, x_prev) per step."""
    ts, table = ddim_schedule(cfg, S, eta)
    img = x_T
    n = len(ts)
    for i, step in enumerate(np.flip(ts)):
        index = n - i - 1
        t = torch.full((img.shape[0],), int(step), dtype=torch.long)
        if mask is not None:   # inpainting blend, ddim.py:146-149 (q_noise[i]: the randn_like draw of q_sample)
            img = q_sample(cfg, x0, t, q_noise[i]) * mask + (1.0 - mask) * img
        e_t = guided_eps(sd, cfg, img, t, cond, unconditional_conditioning, unconditional_guidance_scale)
        nz = None if noise is None else noise[i]
        qf = (lambda p: vq_quantize(p, sd[AE_PREFIX + "quantize.embedding.weight"])[0]) if quantize_x0 else None
        x_prev, pred_x0 = ddim_step(img, e_t, table[index], nz, temperature, quantize=qf)
        if record is not None:
            record.append((img, t, e_t, pred_x0, x_prev))
        img = x_prev
    return img


# --------------------------------------------------------------------------------------
# ancestral DDPM sampling (ddpm.py)
# --------------------------------------------------------------------------------------


def _extract(a, t, x_shape):
    """extract_into_tensor, lidm/modules/basic.py:219-222."""
    b = t.shape[0]
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


def ddpm_p_sample(sched, x, eps, t, noise, temperature=1.0, clip_denoised=False):
    """LatentDiffusion.p_sample given the model output (ddpm.py:1090-1119 with p_mean_variance :1059-1088,
    predict_start_from_noise :219-223, q_posterior :225-232), eps parameterisation."""
    x_recon = _extract(sched["sqrt_recip_alphas_cumprod"], t, x.shape) * x - \
        _extract(sched["sqrt_recipm1_alphas_cumprod"], t, x.shape) * eps
    if clip_denoised:
        x_recon = x_recon.clamp(-1.0, 1.0)
    mean = _extract(sched["posterior_mean_coef1"], t, x.shape) * x_recon + \
        _extract(sched["posterior_mean_coef2"], t, x.shape) * x
    logvar = _extract(sched["posterior_log_variance_clipped"], t, x.shape)
    nz = noise * temperature
    nonzero_mask = (1 - (t == 0).float()).reshape(x.shape[0], *((1,) * (len(x.shape) - 1)))
    return mean + nonzero_mask * (0.5 * logvar).exp() * nz


@torch.no_grad()
def ddpm_sample(sd, cfg: LidmConfig, x_T, timesteps, noise, cond=None, record=None):
    """LatentDiffusion.sample -> p_sample_loop (ddpm.py:1228-1244, 1177-1226): ancestral sampling over
    t = timesteps-1 .. 0 with pre-generated per-step noise (noise[i] for the i-th iteration)."""
    sched = register_schedule(cfg)
    img = x_T
    for i, tv in enumerate(reversed(range(0, timesteps))):
        t = torch.full((img.shape[0],), tv, dtype=torch.long)
        eps = apply_model(sd, cfg, img, t, cond)
        if record is not None:
            record.append((img, t, eps))
        img = ddpm_p_sample(sched, img, eps, t, noise[i])
    return img


# --------------------------------------------------------------------------------------
# back-projection (lidar_utils.py, numpy)
# --------------------------------------------------------------------------------------


def custom_to_unit(x: np.ndarray) -> np.ndarray:
    """scripts/sample.py:29-32: (clip(x,-1,1)+1)/2 in the array's dtype (float32)."""
    return (np.clip(x, -1.0, 1.0) + 1.0) / 2.0


def range2pcd(range_img, fov, depth_range, depth_scale, log_scale=True, **kwargs):
    """lidm/utils/lidar_utils.py:134-172 (without label/color plumbing)."""
    size = range_img.shape
    fov_up = fov[0] / 180.0 * np.pi
    fov_down = fov[1] / 180.0 * np.pi
    fov_range = abs(fov_down) + abs(fov_up)
    depth = (range_img * depth_scale).flatten()
    if log_scale:
        depth = np.exp2(depth) - 1
    scan_x, scan_y = np.meshgrid(np.arange(size[1]), np.arange(size[0]))
    scan_x = scan_x.astype(np.float64) / size[1]
    scan_y = scan_y.astype(np.float64) / size[0]
    yaw = (np.pi * (scan_x * 2 - 1)).flatten()
    pitch = ((1.0 - scan_y) * fov_range - abs(fov_down)).flatten()
    pcd = np.zeros((len(yaw), 3))
    pcd[:, 0] = np.cos(yaw) * np.cos(pitch) * depth
    pcd[:, 1] = -np.sin(yaw) * np.cos(pitch) * depth
    pcd[:, 2] = np.sin(pitch) * depth
    mask = np.logical_and(depth > depth_range[0], depth < depth_range[1])
    return pcd[mask, :], mask


def range2xyz(range_img, fov, depth_range, depth_scale, log_scale=True, **kwargs):
    """lidm/utils/lidar_utils.py:175-204."""
    size = range_img.shape
    fov_up = fov[0] / 180.0 * np.pi
    fov_down = fov[1] / 180.0 * np.pi
    fov_range = abs(fov_down) + abs(fov_up)
    depth = (np.exp2(range_img * depth_scale) - 1) if log_scale else range_img
    scan_x, scan_y = np.meshgrid(np.arange(size[1]), np.arange(size[0]))
    scan_x = scan_x.astype(np.float64) / size[1]
    scan_y = scan_y.astype(np.float64) / size[0]
    yaw = np.pi * (scan_x * 2 - 1)
    pitch = (1.0 - scan_y) * fov_range - abs(fov_down)
    xyz = -np.ones((3, *size))
    xyz[0] = np.cos(yaw) * np.cos(pitch) * depth
    xyz[1] = -np.sin(yaw) * np.cos(pitch) * depth
    xyz[2] = np.sin(pitch) * depth
    mask = np.logical_and(depth > depth_range[0], depth < depth_range[1])
    xyz[:, ~mask] = -1
    return xyz


def rel_l2(a, b) -> float:
    a = torch.as_tensor(a).double().flatten()
    b = torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
