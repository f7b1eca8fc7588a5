"""TEST INFRASTRUCTURE ONLY (imported by tests/ only).  CPU restatement of the condition encoder of the layout-conditioned
LiDM ("next"): `LayoutTransformerEncoder.forward` (reference
lidm/modules/encoders/layout_encoder.py:222-281) with its `Transformer` / `ResidualAttentionBlock` /
`QKVMultiheadAttention` / `MLP` (layout_encoder.py:32-137), as plain functions over the module's state dict.
Pinned bit-exactly against the unmodified reference module on CPU (tests/test_oracle_layout.py, fixture
tests/golden/layout_encoder.npz written by oracle/make_golden_layout.py).  No CUDA path consumes it yet: it is the
first gate for the layout U-Net row of the next round."""
import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _ln(sd, prefix, x):
    # LayerNorm subclass of the reference: computes in fp32, returns the input dtype (layout_encoder.py:23-29)
    return F.layer_norm(x.float(), (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5).to(x.dtype)


def qkv_multihead_attention(qkv: torch.Tensor, n_heads: int, key_padding_mask=None) -> torch.Tensor:
    """QKVMultiheadAttention.forward (layout_encoder.py:65-84): qkv (B, L, heads*3*ch), per head [q | k | v]."""
    bs, n_ctx, width = qkv.shape
    ch = width // n_heads // 3
    scale = 1 / math.sqrt(math.sqrt(ch))
    q, k, v = torch.split(qkv.view(bs, n_ctx, n_heads, -1), ch, dim=-1)
    w = torch.einsum("bthc,bshc->bhts", q * scale, k * scale)
    if key_padding_mask is not None:
        w = w.masked_fill(key_padding_mask.unsqueeze(1).unsqueeze(2), float("-inf"))
    w = torch.softmax(w.float(), dim=-1).type(w.dtype)
    return torch.einsum("bhts,bshc->bthc", w, v).reshape(bs, n_ctx, -1)


def residual_attention_block(sd, prefix: str, x: torch.Tensor, n_heads: int, key_padding_mask=None) -> torch.Tensor:
    """ResidualAttentionBlock.forward (layout_encoder.py:105-108)."""
    h = _lin(sd, prefix + ".attn.c_qkv", _ln(sd, prefix + ".ln_1", x))
    x = x + _lin(sd, prefix + ".attn.c_proj", qkv_multihead_attention(h, n_heads, key_padding_mask))
    h = F.gelu(_lin(sd, prefix + ".mlp.c_fc", _ln(sd, prefix + ".ln_2", x)))
    return x + _lin(sd, prefix + ".mlp.c_proj", h)


def patch_boxes(feature_map_size: Sequence[int], rows: int) -> torch.Tensor:
    """The (x0, y0, x1, y1) unit boxes of the image patches at one attention resolution (layout_encoder.py:198-204):
    `rows` patch rows, columns scaled by the feature map's aspect ratio, row-major."""
    cols = int(feature_map_size[1] / (feature_map_size[0] / rows))
    di, dj = 1.0 / rows, 1.0 / cols
    return torch.FloatTensor([(dj * j, di * i, dj * (j + 1), di * (i + 1)) for i in range(rows) for j in range(cols)])


def layout_encoder_forward(sd: Dict[str, torch.Tensor], layout: torch.Tensor, *, num_layers: int, num_heads: int,
                           used_condition_types: Sequence[str], feature_map_size: Sequence[int] = (8, 128),
                           resolution_to_attention: Sequence[int] = (), use_positional_embedding: bool = False,
                           use_final_ln: bool = True, use_key_padding_mask: bool = False,
                           not_use_layout_fusion_module: bool = False, obj_mask=None) -> Dict[str, torch.Tensor]:
    """LayoutTransformerEncoder.forward (layout_encoder.py:222-281).  layout (B, L, 13) = [bbox 8 | bbox_2d 4 | class 1]."""
    out: Dict[str, torch.Tensor] = {}
    obj_bbox, obj_bbox_2d, obj_class = torch.split(layout, [8, 4, 1], dim=-1)
    is_valid_obj = obj_class > 0
    obj_class = obj_class.squeeze(dim=-1)
    xf_in = sd["positional_embedding"][None] if use_positional_embedding else None

    def acc(term):
        return term if xf_in is None else xf_in + term

    if "obj_class" in used_condition_types:
        e = F.embedding(obj_class.long(), sd["obj_class_embedding.weight"])
        xf_in = acc(e)
        out["obj_class_embedding"] = e.permute(0, 2, 1)
    if "obj_bbox" in used_condition_types:
        e = _lin(sd, "obj_bbox_embedding", obj_bbox_2d.float())
        enc = _lin(sd, "obj_bbox_encoding", obj_bbox.float())
        xf_in = (e + enc) if xf_in is None else (xf_in + e + enc)
        out["obj_bbox_embedding"] = e.permute(0, 2, 1)
        for r in resolution_to_attention:
            pe = _lin(sd, "obj_bbox_embedding", patch_boxes(feature_map_size, r)).unsqueeze(0)
            out[f"image_patch_bbox_embedding_for_resolution{r}"] = torch.repeat_interleave(pe, e.shape[0], dim=0).permute(0, 2, 1)
    if "obj_mask" in used_condition_types:
        xf_in = acc(_lin(sd, "obj_mask_embedding", obj_mask.view(*obj_mask.shape[:2], -1).float()))
    if "is_valid_obj" in used_condition_types:
        out["key_padding_mask"] = (1 - is_valid_obj.int()).bool()
    kpm = out["key_padding_mask"] if use_key_padding_mask else None
    x = xf_in.float()
    if not not_use_layout_fusion_module:
        for i in range(num_layers):
            x = residual_attention_block(sd, f"transform.resblocks.{i}", x, num_heads, kpm)
    if use_final_ln:
        x = _ln(sd, "final_ln", x)
    out["xf_proj"] = _lin(sd, "transformer_proj", x[:, 0])
    out["xf_out"] = x.permute(0, 2, 1)
    return out


def synthetic_layout(B: int, L: int, n_classes: int, seed: int) -> torch.Tensor:
    """A (B, L, 13) layout tensor with the reference's field order; trailing objects of each sample are padding (class 0)."""
    g = torch.Generator().manual_seed(seed)
    box = torch.rand(B, L, 8, generator=g) * 2 - 1
    box2d = torch.rand(B, L, 4, generator=g)
    cls = torch.randint(1, n_classes, (B, L, 1), generator=g).float()
    for b in range(B):
        cls[b, L - 1 - (b % 3):] = 0
    return torch.cat([box, box2d, cls], dim=-1)


# --------------------------------------------------------------------------------------------------------------------
# ObjectAwareCrossAttention (reference lidm/modules/unets/object_cross_unet.py:380-565): every image token attends to
# the image tokens AND the layout tokens; queries / keys are [content | positional] halves (head width 2 x C / heads),
# values are content only.

def _gn32(sd, prefix, x):
    # normalization() = GroupNorm32(32, C) computing in fp32 (lidm/modules/unets/nn.py:17-19,93-100)
    return F.group_norm(x.float(), 32, sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-5).type(x.dtype)


def _conv1(sd, prefix, x):
    return F.conv1d(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def object_aware_cross_attention(sd: Dict[str, torch.Tensor], x: torch.Tensor, cond: Dict[str, torch.Tensor], *,
                                 num_heads: int, resolution_rows: int, pos_scale: float = 1.0, norm_first: bool = False,
                                 norm_for_obj_embedding: bool = False) -> torch.Tensor:
    """ObjectAwareCrossAttention.forward (object_cross_unet.py:447-565, use_key_padding_mask False as shipped).
    x (B, C, H, W); cond = outputs of the layout encoder.  Returns x + proj_out(attention)."""
    b, c, *spatial = x.shape
    x3 = x.reshape(b, c, -1)
    qkv = _conv1(sd, "qkv_projector", _gn32(sd, "norm_for_qkv", x3))
    C, L1 = c, qkv.shape[2]
    L2 = cond["obj_bbox_embedding"].shape[-1]
    cp = int(C * pos_scale)

    def positional(src, norm_name):
        if norm_first:
            return _conv1(sd, "layout_position_embedding_projector", _gn32(sd, norm_name, src))
        return _gn32(sd, norm_name, _conv1(sd, "layout_position_embedding_projector", src))

    img_pos = positional(cond[f"image_patch_bbox_embedding_for_resolution{resolution_rows}"],
                         "norm_for_image_patch_positional_embedding").reshape(b * num_heads, cp // num_heads, L1)
    q_c, k_c, v_c = (t.reshape(b * num_heads, C // num_heads, L1) for t in qkv.split(C, dim=1))
    q_img = torch.cat([q_c, img_pos], dim=1)
    k_img = torch.cat([k_c, img_pos], dim=1)
    lay_pos = positional(cond["obj_bbox_embedding"], "norm_for_layout_positional_embedding").reshape(
        b * num_heads, cp // num_heads, L2)
    xf = _gn32(sd, "norm_for_obj_embedding", cond["xf_out"]) if norm_for_obj_embedding else cond["xf_out"]
    content = (xf + _gn32(sd, "norm_for_obj_class_embedding", cond["obj_class_embedding"])) / 2
    k_l, v_l = _conv1(sd, "layout_content_embedding_projector", content).split(C, dim=1)
    k_lay = torch.cat([k_l.reshape(b * num_heads, C // num_heads, L2), lay_pos], dim=1)
    v_lay = v_l.reshape(b * num_heads, C // num_heads, L2)
    k_mix = torch.cat([k_img, k_lay], dim=2)
    v_mix = torch.cat([v_c, v_lay], dim=2)
    scale = 1 / math.sqrt(math.sqrt(int((1 + pos_scale) * C) // num_heads))
    w = torch.einsum("bct,bcs->bts", q_img * scale, k_mix * scale)
    w = torch.softmax(w.float(), dim=-1).type(w.dtype)
    a = torch.einsum("bts,bcs->bct", w, v_mix).reshape(b, C, L1)
    return (x3 + _conv1(sd, "proj_out", a)).reshape(b, c, *spatial)


# --------------------------------------------------------------------------------------------------------------------
# LayoutDiffusionUNetModel (reference lidm/modules/unets/object_cross_unet.py:632-952): guided-diffusion style U-Net with
# FiLM (scale-shift) ResBlocks, ResBlock up/down-sampling, ZERO-padded 3x3 convs and ObjectAwareCrossAttention blocks.

def _silu(x):
    return x * torch.sigmoid(x)        # SiLU of the reference (object_cross_unet.py:44-47)


def _conv2(sd, prefix, x, padding):
    return F.conv2d(x, sd[prefix + ".weight"], sd[prefix + ".bias"], padding=padding)


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
    """lidm/modules/unets/nn.py:103-121: [cos | sin] halves."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def layout_res_block(sd, prefix: str, x: torch.Tensor, emb: torch.Tensor, *, up: bool = False, down: bool = False,
                     use_scale_shift_norm: bool = True) -> torch.Tensor:
    """ResBlock.forward (object_cross_unet.py:253-283); Upsample / Downsample without conv (112-171)."""
    h = _silu(_gn32(sd, prefix + ".in_layers.0", x))
    if up:
        h, x = (F.interpolate(t, scale_factor=2, mode="nearest") for t in (h, x))
    elif down:
        h, x = (F.avg_pool2d(t, kernel_size=2, stride=2) for t in (h, x))
    h = _conv2(sd, prefix + ".in_layers.2", h, 1)
    e = F.linear(_silu(emb), sd[prefix + ".emb_layers.1.weight"], sd[prefix + ".emb_layers.1.bias"]).type(h.dtype)
    e = e[..., None, None]
    if use_scale_shift_norm:
        scale, shift = torch.chunk(e, 2, dim=1)
        h = _gn32(sd, prefix + ".out_layers.0", h) * (1 + scale) + shift
        h = _conv2(sd, prefix + ".out_layers.3", _silu(h), 1)
    else:
        h = _conv2(sd, prefix + ".out_layers.3", _silu(_gn32(sd, prefix + ".out_layers.0", h + e)), 1)
    if prefix + ".skip_connection.weight" in sd:
        w = sd[prefix + ".skip_connection.weight"]
        x = F.conv2d(x, w, sd[prefix + ".skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


def layout_unet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, timesteps: torch.Tensor,
                        layout_outputs: Dict[str, torch.Tensor], *, model_channels: int, channel_mult: Sequence[int],
                        num_res_blocks: int, attention_ds: Sequence[int], image_size: Sequence[int], num_head_channels: int,
                        num_attention_blocks: int = 1, use_scale_shift_norm: bool = True, pos_scale: float = 1.0,
                        norm_first: bool = False, norm_for_obj_embedding: bool = False) -> torch.Tensor:
    """LayoutDiffusionUNetModel.forward (object_cross_unet.py:923-951) for resblock_updown=True (the shipped setting);
    the module list is walked exactly as the constructor builds it (object_cross_unet.py:742-912)."""
    emb = F.linear(timestep_embedding(timesteps, model_channels), sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(_silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    emb = emb + layout_outputs["xf_proj"].to(emb)

    def attn(prefix, h, ds):
        ch = h.shape[1]
        return object_aware_cross_attention(
            {k[len(prefix) + 1:]: v for k, v in sd.items() if k.startswith(prefix + ".")}, h, layout_outputs,
            num_heads=ch // num_head_channels, resolution_rows=int(image_size[0] // ds), pos_scale=pos_scale,
            norm_first=norm_first, norm_for_obj_embedding=norm_for_obj_embedding)

    res = lambda prefix, h, **kw: layout_res_block(sd, prefix, h, emb, use_scale_shift_norm=use_scale_shift_norm, **kw)
    hs: List[torch.Tensor] = []
    h = _conv2(sd, "input_blocks.0.0", x, 1)
    hs.append(h)
    idx, ds = 1, 1
    for level, _ in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            h = res(f"input_blocks.{idx}.0", h)
            if ds in attention_ds:
                for a in range(num_attention_blocks):
                    h = attn(f"input_blocks.{idx}.{1 + a}", h, ds)
            hs.append(h)
            idx += 1
        if level != len(channel_mult) - 1:
            h = res(f"input_blocks.{idx}.0", h, down=True)
            hs.append(h)
            idx += 1
            ds *= 2
    h = res("middle_block.0", h)
    h = attn("middle_block.1", h, ds)
    h = res("middle_block.2", h)
    idx = 0
    for level in reversed(range(len(channel_mult))):
        for i in range(num_res_blocks + 1):
            h = res(f"output_blocks.{idx}.0", torch.cat([h, hs.pop()], dim=1))
            j = 1
            if ds in attention_ds:
                for _ in range(num_attention_blocks):
                    h = attn(f"output_blocks.{idx}.{j}", h, ds)
                    j += 1
            if level and i == num_res_blocks:
                h = res(f"output_blocks.{idx}.{j}", h, up=True)
                ds //= 2
            idx += 1
    return _conv2(sd, "out.2", _silu(_gn32(sd, "out.0", h)), 1)


def seeded_state_dict(shapes: Dict[str, Sequence[int]], seed: int, std: float) -> Dict[str, torch.Tensor]:
    """Deterministic fp32 weights for a {name: shape} table, independent of torch's RNG stream and of module
    construction order: every tensor is N(0, std) from a numpy generator keyed by (seed, crc32(name)).  Fixtures store
    the shape table only; the generator script loads these weights into the reference module, tests feed the oracle."""
    import zlib
    import numpy as np
    out = {}
    for name in sorted(shapes):
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        out[name] = torch.from_numpy((rng.standard_normal(tuple(shapes[name])) * std).astype(np.float32))
    return out
