"""State-dict key/shape specification of the reference LiDM and a seeded synthetic state-dict.

Checkpoints are not available offline, so tests/bench use random-init weights.  The reference's own
random init is degenerate for parity testing (34 zero-init tensors make
eps == 0, the U(+-1/16384) codebook makes the quantised decode insensitive to z), therefore this
module generates a *platform-independent* (numpy PCG64) state-dict in the reference's own key scheme
in which every tensor is non-degenerate.  `oracle/make_golden.py` loads it with
`load_state_dict(strict=True)` into the real reference modules, which pins names and shapes.

Key scheme (reference): `model.diffusion_model.*` (lidm/models/diffusion/ddpm.py:91, 2306-2311),
`first_stage_model.{decoder,quantize.embedding,post_quant_conv}.*` (lidm/models/ae/autoencoder.py:42-50).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

from .config import AEConfig, LidmConfig, UNetConfig

UNET_PREFIX = "model.diffusion_model."
AE_PREFIX = "first_stage_model."


def _conv(spec, name, cout, cin, kh, kw, zero_init=False):
    spec[name + ".weight"] = ((cout, cin, kh, kw), "conv_zero" if zero_init else "conv")
    spec[name + ".bias"] = ((cout,), "bias")


def _conv1d(spec, name, cout, cin, zero_init=False):
    spec[name + ".weight"] = ((cout, cin, 1), "conv_zero" if zero_init else "conv")
    spec[name + ".bias"] = ((cout,), "bias")


def _linear(spec, name, cout, cin):
    spec[name + ".weight"] = ((cout, cin), "conv")
    spec[name + ".bias"] = ((cout,), "bias")


def _norm(spec, name, c):
    spec[name + ".weight"] = ((c,), "gamma")
    spec[name + ".bias"] = ((c,), "beta")


def unet_blocks(cfg: UNetConfig):
    """Walk the U-Net topology exactly like UNetModel.__init__ (reference openaimodel.py:509-687).

    Returns (input_blocks, middle, output_blocks, final_ch) where every block is a list of layer tuples:
      ("conv", cin, cout)              3x3 circular conv (input_blocks.0)
      ("res", cin, cout)               ResBlock
      ("attn", ch, heads)              AttentionBlock
      ("st", ch, heads, depth, ctx)    SpatialTransformer (use_spatial_transformer; lidm/modules/attention.py:218-261)
      ("down", ch)                     Downsample (conv 3x3 stride 2)
      ("up", ch)                       Upsample (nearest x2 + conv 3x3)
    """
    if cfg.use_scale_shift_norm or cfg.resblock_updown:
        raise NotImplementedError("use_scale_shift_norm / resblock_updown U-Nets are not on the B200 path")
    if cfg.use_spatial_transformer and cfg.context_dim is None:
        raise ValueError("use_spatial_transformer needs context_dim (openaimodel.py:474-475)")

    def attn_layer(ch):
        heads = ch // cfg.num_head_channels
        if cfg.use_spatial_transformer:
            return ("st", ch, heads, cfg.transformer_depth, cfg.context_dim)
        return ("attn", ch, heads)
    mc = cfg.model_channels
    inputs = [[("conv", cfg.in_channels, mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", ch, mult * mc)]
            ch = mult * mc
            if ds in cfg.attention_resolutions:
                layers.append(attn_layer(ch))
            inputs.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            inputs.append([("down", ch)])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch), attn_layer(ch), ("res", ch, ch)]
    outputs = []
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * mult)]
            ch = mc * mult
            if ds in cfg.attention_resolutions:
                layers.append(attn_layer(ch))
            if level and i == cfg.num_res_blocks:
                layers.append(("up", ch))
                ds //= 2
            outputs.append(layers)
    return inputs, middle, outputs, ch


def _block_spec(spec, prefix, layers, ted):
    for j, layer in enumerate(layers):
        p = f"{prefix}.{j}"
        kind = layer[0]
        if kind == "conv":
            _conv(spec, p, layer[2], layer[1], 3, 3)
        elif kind == "res":
            _, cin, cout = layer
            _norm(spec, p + ".in_layers.0", cin)
            _conv(spec, p + ".in_layers.2", cout, cin, 3, 3)
            _linear(spec, p + ".emb_layers.1", cout, ted)
            _norm(spec, p + ".out_layers.0", cout)
            _conv(spec, p + ".out_layers.3", cout, cout, 3, 3, zero_init=True)
            if cin != cout:
                _conv(spec, p + ".skip_connection", cout, cin, 1, 1)
        elif kind == "attn":
            ch = layer[1]
            _norm(spec, p + ".norm", ch)
            _conv1d(spec, p + ".qkv", 3 * ch, ch)
            _conv1d(spec, p + ".proj_out", ch, ch, zero_init=True)
        elif kind == "st":
            _, ch, heads, depth, ctx = layer
            _norm(spec, p + ".norm", ch)
            _conv(spec, p + ".proj_in", ch, ch, 1, 1)
            for d in range(depth):
                b = f"{p}.transformer_blocks.{d}"
                for a, kdim in (("attn1", ch), ("attn2", ctx)):
                    spec[f"{b}.{a}.to_q.weight"] = ((ch, ch), "conv")
                    spec[f"{b}.{a}.to_k.weight"] = ((ch, kdim), "conv")
                    spec[f"{b}.{a}.to_v.weight"] = ((ch, kdim), "conv")
                    _linear(spec, f"{b}.{a}.to_out.0", ch, ch)
                _linear(spec, f"{b}.ff.net.0.proj", 8 * ch, ch)
                _linear(spec, f"{b}.ff.net.2", ch, 4 * ch)
                for n in ("norm1", "norm2", "norm3"):
                    _norm(spec, f"{b}.{n}", ch)
            _conv(spec, p + ".proj_out", ch, ch, 1, 1, zero_init=True)
        elif kind == "down":
            _conv(spec, p + ".op", layer[1], layer[1], 3, 3)
        elif kind == "up":
            _conv(spec, p + ".conv", layer[1], layer[1], 3, 3)
        else:  # pragma: no cover
            raise ValueError(kind)


def layout_unet_blocks(cfg: UNetConfig):
    """Walk LayoutDiffusionUNetModel.__init__ (reference lidm/modules/unets/object_cross_unet.py:742-912,
    resblock_updown = True).  Layer tuples: ("conv", cin, cout), ("fres", cin, cout, updown) with updown in
    {None, "up", "down"}, ("oaca", ch, ds)."""
    mc = cfg.model_channels
    inputs = [[("conv", cfg.in_channels, mc)]]
    chans = [mc]
    ch, ds = mc, 1
    attn = lambda c, d: [("oaca", c, d)] * cfg.num_attention_blocks if d in cfg.attention_resolutions else []
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("fres", ch, mult * mc, None)]
            ch = mult * mc
            layers += attn(ch, ds)
            inputs.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            inputs.append([("fres", ch, ch, "down")])
            chans.append(ch)
            ds *= 2
    middle = [("fres", ch, ch, None), ("oaca", ch, ds), ("fres", ch, ch, None)]
    outputs = []
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            layers = [("fres", ch + ich, mc * mult, None)]
            ch = mc * mult
            layers += attn(ch, ds)
            if level and i == cfg.num_res_blocks:
                layers.append(("fres", ch, ch, "up"))
                ds //= 2
            outputs.append(layers)
    return inputs, middle, outputs, ch


def _layout_block_spec(spec, prefix, layers, ted, enc_ch):
    for j, layer in enumerate(layers):
        p = f"{prefix}.{j}"
        if layer[0] == "conv":
            _conv(spec, p, layer[2], layer[1], 3, 3)
        elif layer[0] == "fres":       # ResBlock with use_scale_shift_norm (object_cross_unet.py:176-251)
            _, cin, cout, _ = layer
            _norm(spec, p + ".in_layers.0", cin)
            _conv(spec, p + ".in_layers.2", cout, cin, 3, 3)
            _linear(spec, p + ".emb_layers.1", 2 * cout, ted)
            _norm(spec, p + ".out_layers.0", cout)
            _conv(spec, p + ".out_layers.3", cout, cout, 3, 3, zero_init=True)
            if cin != cout:
                _conv(spec, p + ".skip_connection", cout, cin, 1, 1)
        elif layer[0] == "oaca":       # ObjectAwareCrossAttention (object_cross_unet.py:430-446, norm_first False)
            ch = layer[1]
            _conv1d(spec, p + ".qkv_projector", 3 * ch, ch)
            _norm(spec, p + ".norm_for_qkv", ch)
            _conv1d(spec, p + ".layout_content_embedding_projector", 2 * ch, enc_ch)
            _conv1d(spec, p + ".layout_position_embedding_projector", ch, enc_ch)
            _norm(spec, p + ".norm_for_obj_class_embedding", enc_ch)
            _norm(spec, p + ".norm_for_layout_positional_embedding", ch)
            _norm(spec, p + ".norm_for_image_patch_positional_embedding", ch)
            _conv1d(spec, p + ".proj_out", ch, ch, zero_init=True)
        else:  # pragma: no cover
            raise ValueError(layer[0])


def layout_unet_param_spec(cfg: UNetConfig, prefix: str = UNET_PREFIX) -> "OrderedDict[str, Tuple[tuple, str]]":
    spec: OrderedDict = OrderedDict()
    mc, ted = cfg.model_channels, cfg.time_embed_dim
    _linear(spec, prefix + "time_embed.0", ted, mc)
    _linear(spec, prefix + "time_embed.2", ted, ted)
    inputs, middle, outputs, ch = layout_unet_blocks(cfg)
    for i, layers in enumerate(inputs):
        _layout_block_spec(spec, f"{prefix}input_blocks.{i}", layers, ted, cfg.encoder_channels)
    _layout_block_spec(spec, f"{prefix}middle_block", middle, ted, cfg.encoder_channels)
    for i, layers in enumerate(outputs):
        _layout_block_spec(spec, f"{prefix}output_blocks.{i}", layers, ted, cfg.encoder_channels)
    _norm(spec, prefix + "out.0", ch)
    _conv(spec, prefix + "out.2", cfg.out_channels, mc, 3, 3, zero_init=True)
    return spec


def efficient_unet_param_spec(cfg: UNetConfig, prefix: str = UNET_PREFIX) -> "OrderedDict[str, Tuple[tuple, str]]":
    """EfficientUNet tensors (reference lidm/modules/unets/efficient_unet.py:188-260, coords_encoding 'fourier_features':
    the `coords` / `coords_encoding.*` / `*.scale` buffers are constants and not part of the spec)."""
    spec: OrderedDict = OrderedDict()
    mc, ted = cfg.model_channels, cfg.model_channels * 4
    H, W = cfg.image_size
    extra = 2 * (int(np.ceil(np.log2(H))) + int(np.ceil(np.log2(W))))
    _linear(spec, prefix + "time_embedding.1", ted, mc)
    _linear(spec, prefix + "time_embedding.3", ted, ted)
    C = [mc] + [mc * m for m in cfg.channel_mult]
    N = list(cfg.num_residual_blocks)
    _conv(spec, prefix + "in_conv", C[0], cfg.in_channels + extra, 3, 3)

    def block(p, cin, cout, n, down=False, up=False, attn=False):
        if down:
            _conv(spec, p + ".downsample.0", cout, cin, 3, 3)
        for i in range(n):
            rp = f"{p}.residual_blocks.{i}"
            ci = cout if (i != 0 or down) else cin
            _norm(spec, rp + ".norm1", ci)
            _conv(spec, rp + ".conv1", cout, ci, 3, 3)
            _linear(spec, rp + ".norm2.proj.1", 2 * cout, ted)
            _conv(spec, rp + ".conv2", cout, cout, 3, 3, zero_init=True)
            if ci != cout:
                _conv(spec, rp + ".skip", cout, ci, 1, 1)
        if attn:
            ap = p + ".self_attn_block"
            _norm(spec, ap + ".norm", cout)
            spec[ap + ".attn.in_proj_weight"] = ((3 * cout, cout), "conv")
            spec[ap + ".attn.in_proj_bias"] = ((3 * cout,), "bias")
            spec[ap + ".attn.out_proj.weight"] = ((cout, cout), "conv_zero")
            spec[ap + ".attn.out_proj.bias"] = ((cout,), "bias")
        if up:
            _conv(spec, p + ".upsample.1", cout, cout, 3, 3)

    block(prefix + "d_block1", C[0], C[1], N[0])
    block(prefix + "d_block2", C[1], C[2], N[1], down=True)
    block(prefix + "d_block3", C[2], C[3], N[2], down=True)
    block(prefix + "d_block4", C[3], C[4], N[3], down=True, attn=True)
    block(prefix + "u_block4", C[4], C[3], N[3], up=True, attn=True)
    block(prefix + "u_block3", 2 * C[3], C[2], N[2], up=True)
    block(prefix + "u_block2", 2 * C[2], C[1], N[1], up=True)
    block(prefix + "u_block1", 2 * C[1], C[0], N[0])
    _conv(spec, prefix + "out_conv", cfg.out_channels, C[0], 3, 3, zero_init=True)
    return spec


def unet_param_spec(cfg: UNetConfig, prefix: str = UNET_PREFIX) -> "OrderedDict[str, Tuple[tuple, str]]":
    if cfg.unet_type == "layout":
        return layout_unet_param_spec(cfg, prefix)
    if cfg.unet_type == "efficient":
        return efficient_unet_param_spec(cfg, prefix)
    spec: OrderedDict = OrderedDict()
    mc, ted = cfg.model_channels, cfg.time_embed_dim
    _linear(spec, prefix + "time_embed.0", ted, mc)
    _linear(spec, prefix + "time_embed.2", ted, ted)
    inputs, middle, outputs, ch = unet_blocks(cfg)
    for i, layers in enumerate(inputs):
        _block_spec(spec, f"{prefix}input_blocks.{i}", layers, ted)
    _block_spec(spec, f"{prefix}middle_block", middle, ted)
    for i, layers in enumerate(outputs):
        _block_spec(spec, f"{prefix}output_blocks.{i}", layers, ted)
    _norm(spec, prefix + "out.0", ch)
    _conv(spec, prefix + "out.2", cfg.out_channels, mc, 3, 3, zero_init=True)
    return spec


def decoder_levels(cfg: AEConfig):
    """Decoder topology (reference model_lidm.py:315-383): list over i_level (reversed order of execution)
    of dicts {blocks:[(cin,cout)], kernel:(kh,kw), stride:None|(sh,sw)}; plus block_in at the lowest res."""
    stride2kernel = {(2, 2): (3, 3), (1, 2): (1, 4)}
    nres = len(cfg.ch_mult)
    block_in = cfg.ch * cfg.ch_mult[nres - 1]
    top = block_in
    levels = {}
    for i_level in reversed(range(nres)):
        stride = tuple(cfg.strides[i_level - 1]) if i_level > 0 else None
        kernel = stride2kernel[stride] if stride is not None else (1, 4)
        block_out = cfg.ch * cfg.ch_mult[i_level]
        blocks = []
        for _ in range(cfg.num_res_blocks + 1):
            blocks.append((block_in, block_out))
            block_in = block_out
        levels[i_level] = dict(blocks=blocks, kernel=kernel, stride=stride, ch=block_in)
    return top, levels, block_in


def _resnet_spec(spec, p, cin, cout, k):
    _norm(spec, p + ".norm1", cin)
    _conv(spec, p + ".conv1", cout, cin, k[0], k[1])
    _norm(spec, p + ".norm2", cout)
    _conv(spec, p + ".conv2", cout, cout, k[0], k[1])
    if cin != cout:
        _conv(spec, p + ".nin_shortcut", cout, cin, 1, 1)


UPSAMPLE_STRIDE2KERNEL = {(1, 2): (1, 5), (1, 4): (1, 7), (2, 1): (5, 1), (2, 2): (3, 3)}


def ae_param_spec(cfg: AEConfig, prefix: str = AE_PREFIX) -> "OrderedDict[str, Tuple[tuple, str]]":
    """Decode-side tensors of VQModelInterface (encoder tensors are not on the path)."""
    if cfg.attn_levels:
        raise NotImplementedError("decoder level attention is not used by the named configs")
    spec: OrderedDict = OrderedDict()
    spec[prefix + "quantize.embedding.weight"] = ((cfg.n_embed, cfg.embed_dim), "codebook")
    _conv(spec, prefix + "post_quant_conv", cfg.z_channels, cfg.embed_dim, 1, 1)
    d = prefix + "decoder."
    top, levels, last = decoder_levels(cfg)
    _conv(spec, d + "conv_in", top, cfg.z_channels, 3, 3)
    _resnet_spec(spec, d + "mid.block_1", top, top, (3, 3))
    _norm(spec, d + "mid.attn_1.norm", top)
    for n in ("q", "k", "v", "proj_out"):
        _conv(spec, d + f"mid.attn_1.{n}", top, top, 1, 1)
    _resnet_spec(spec, d + "mid.block_2", top, top, (3, 3))
    for i_level in reversed(range(len(cfg.ch_mult))):
        lv = levels[i_level]
        for i_block, (cin, cout) in enumerate(lv["blocks"]):
            _resnet_spec(spec, d + f"up.{i_level}.block.{i_block}", cin, cout, lv["kernel"])
        if lv["stride"] is not None:
            k = UPSAMPLE_STRIDE2KERNEL[lv["stride"]]
            _conv(spec, d + f"up.{i_level}.upsample.conv", lv["ch"], lv["ch"], k[0], k[1])
    _norm(spec, d + "norm_out", last)
    _conv(spec, d + "conv_out", cfg.out_ch, last, 1, 4)
    return spec


DOWNSAMPLE_STRIDE2KERNEL = {(1, 2): (3, 3), (1, 4): (3, 5), (2, 1): (3, 3), (2, 2): (3, 3)}
DOWNSAMPLE_STRIDE2PAD = {(1, 2): (0, 1, 1, 1), (1, 4): (1, 1, 1, 1), (2, 1): (1, 1, 1, 1), (2, 2): (0, 1, 0, 1)}


def encoder_levels(cfg: AEConfig):
    """Encoder topology (reference model_lidm.py:222-282): list over i_level of dicts
    {blocks:[(cin,cout)], stride:None|(sh,sw), ch}; ResnetBlocks use the default 3x3 kernel."""
    if cfg.attn_levels:
        raise NotImplementedError("encoder level attention is not used by the named configs")
    in_ch_mult = (1,) + tuple(cfg.ch_mult)
    nres = len(cfg.ch_mult)
    levels = []
    block_in = cfg.ch
    for i_level in range(nres):
        block_in = cfg.ch * in_ch_mult[i_level]
        block_out = cfg.ch * cfg.ch_mult[i_level]
        blocks = []
        for _ in range(cfg.num_res_blocks):
            blocks.append((block_in, block_out))
            block_in = block_out
        stride = tuple(cfg.strides[i_level]) if i_level != nres - 1 else None
        levels.append(dict(blocks=blocks, stride=stride, ch=block_in))
    return levels, block_in


def ae_encoder_param_spec(cfg: AEConfig, prefix: str = AE_PREFIX) -> "OrderedDict[str, Tuple[tuple, str]]":
    """Encode-side tensors of VQModelInterface (Encoder + quant_conv; model_lidm.py:222-312, autoencoder.py:42-49)."""
    spec: OrderedDict = OrderedDict()
    e = prefix + "encoder."
    levels, top = encoder_levels(cfg)
    _conv(spec, e + "conv_in", cfg.ch, cfg.in_channels, 3, 3)
    for i_level, lv in enumerate(levels):
        for i_block, (cin, cout) in enumerate(lv["blocks"]):
            _resnet_spec(spec, e + f"down.{i_level}.block.{i_block}", cin, cout, (3, 3))
        if lv["stride"] is not None:
            k = DOWNSAMPLE_STRIDE2KERNEL[lv["stride"]]
            _conv(spec, e + f"down.{i_level}.downsample.conv", lv["ch"], lv["ch"], k[0], k[1])
    _resnet_spec(spec, e + "mid.block_1", top, top, (3, 3))
    _norm(spec, e + "mid.attn_1.norm", top)
    for n in ("q", "k", "v", "proj_out"):
        _conv(spec, e + f"mid.attn_1.{n}", top, top, 1, 1)
    _resnet_spec(spec, e + "mid.block_2", top, top, (3, 3))
    _norm(spec, e + "norm_out", top)
    _conv(spec, e + "conv_out", cfg.z_channels, top, 3, 3)
    _conv(spec, prefix + "quant_conv", cfg.embed_dim, cfg.z_channels, 1, 1)
    return spec


COND_PREFIX = "cond_stage_model."


def layout_encoder_param_spec(le, prefix: str = COND_PREFIX) -> "OrderedDict[str, Tuple[tuple, str]]":
    """LayoutTransformerEncoder tensors (reference lidm/modules/encoders/layout_encoder.py:140-220 with its Transformer,
    :32-137) for the shipped condition types; `le` is a config.LayoutEncoderConfig."""
    spec: OrderedDict = OrderedDict()
    H = le.hidden_dim
    for i in range(le.num_layers):
        p = f"{prefix}transform.resblocks.{i}"
        _linear(spec, p + ".attn.c_qkv", 3 * H, H)
        _linear(spec, p + ".attn.c_proj", H, H)
        _norm(spec, p + ".ln_1", H)
        _linear(spec, p + ".mlp.c_fc", 4 * H, H)
        _linear(spec, p + ".mlp.c_proj", H, 4 * H)
        _norm(spec, p + ".ln_2", H)
    _linear(spec, prefix + "transformer_proj", le.output_dim, H)
    spec[prefix + "obj_class_embedding.weight"] = ((le.num_classes_for_layout_object, H), "codebook")
    _linear(spec, prefix + "obj_bbox_embedding", H, 4)
    _linear(spec, prefix + "obj_bbox_encoding", H, 8)
    if le.use_final_ln:
        _norm(spec, prefix + "final_ln", H)
    return spec


def random_layout_encoder_state_dict(cfg: LidmConfig, seed: int = 0, as_torch: bool = True):
    """Synthetic cond-stage tensors (LayoutTransformerEncoder), drawn from their own generator (the sampling-side
    state-dict of `random_state_dict` does not depend on them)."""
    rng = np.random.Generator(np.random.PCG64(seed + 104729))
    out = OrderedDict((name, _draw(rng, shape, kind, 1.0, 1.0))
                      for name, (shape, kind) in layout_encoder_param_spec(cfg.layout_encoder).items())
    if as_torch:
        import torch
        return OrderedDict((k, torch.from_numpy(v)) for k, v in out.items())
    return out


def param_spec(cfg: LidmConfig):
    spec = unet_param_spec(cfg.unet)
    if cfg.unet.unet_type != "efficient":       # the pixel-space R2DM model has no first stage
        spec.update(ae_param_spec(cfg.ae))
    return spec


def _draw(rng, shape, kind, zero_init_scale, codebook_std):
    n = int(np.prod(shape))
    if kind in ("conv", "conv_zero"):
        fan_in = int(np.prod(shape[1:]))
        b = 1.0 / np.sqrt(fan_in)
        a = (rng.random(n, dtype=np.float32) * 2.0 - 1.0) * np.float32(b)
        if kind == "conv_zero":
            a = a * np.float32(zero_init_scale)
    elif kind == "bias":
        a = (rng.random(n, dtype=np.float32) * 2.0 - 1.0) * np.float32(0.05)
    elif kind == "gamma":
        a = 1.0 + 0.1 * rng.standard_normal(n, dtype=np.float32)
    elif kind == "beta":
        a = 0.1 * rng.standard_normal(n, dtype=np.float32)
    elif kind == "codebook":
        a = np.float32(codebook_std) * rng.standard_normal(n, dtype=np.float32)
    else:  # pragma: no cover
        raise ValueError(kind)
    return np.ascontiguousarray(a.astype(np.float32).reshape(shape))


def random_encoder_state_dict(cfg: LidmConfig, seed: int = 0, as_torch: bool = True):
    """Synthetic encode-side tensors (Encoder + quant_conv), drawn from their own generator so that the sampling-side
    state-dict of `random_state_dict` (and the digest of the committed fixtures) does not depend on them."""
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    out = OrderedDict((name, _draw(rng, shape, kind, 1.0, 1.0)) for name, (shape, kind) in ae_encoder_param_spec(cfg.ae).items())
    if as_torch:
        import torch
        return OrderedDict((k, torch.from_numpy(v)) for k, v in out.items())
    return out


def random_state_dict(cfg: LidmConfig, seed: int = 0, zero_init_scale: float = 1.0,
                      codebook_std: float = 1.0, as_torch: bool = True) -> Dict[str, "np.ndarray"]:
    """Seeded, platform-independent synthetic weights in the reference key scheme.

    conv/linear: U(-b, b), b = 1/sqrt(fan_in)   (PyTorch default init distribution)
    conv_zero  : same distribution x zero_init_scale (the reference zero-inits these; see module docstring)
    bias       : U(-0.05, 0.05)
    gamma/beta : 1 + 0.1 N(0,1) / 0.1 N(0,1)
    codebook   : N(0, codebook_std^2)  (latent scale)
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    out = OrderedDict()
    for name, (shape, kind) in param_spec(cfg).items():
        n = int(np.prod(shape))
        if kind in ("conv", "conv_zero"):
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / np.sqrt(fan_in)
            a = (rng.random(n, dtype=np.float32) * 2.0 - 1.0) * np.float32(b)
            if kind == "conv_zero":
                a = a * np.float32(zero_init_scale)
        elif kind == "bias":
            a = (rng.random(n, dtype=np.float32) * 2.0 - 1.0) * np.float32(0.05)
        elif kind == "gamma":
            a = 1.0 + 0.1 * rng.standard_normal(n, dtype=np.float32)
        elif kind == "beta":
            a = 0.1 * rng.standard_normal(n, dtype=np.float32)
        elif kind == "codebook":
            a = np.float32(codebook_std) * rng.standard_normal(n, dtype=np.float32)
        else:  # pragma: no cover
            raise ValueError(kind)
        out[name] = np.ascontiguousarray(a.astype(np.float32).reshape(shape))
    if as_torch:
        import torch
        return OrderedDict((k, torch.from_numpy(v)) for k, v in out.items())
    return out
