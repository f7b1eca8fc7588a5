"""Configuration for the B200 LiDM sampling path.

Parsed from the reference's own YAML format (`target:` + `params:`), e.g.
`models/lidm/kitti/uncond/config.yaml` (reference), so the drop-in is constructed from the same
file a reference user already has (SURVEY.md section 8(b) "Construction").
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple


@dataclass
class UNetConfig:
    """Mirror of UNetModel.__init__ arguments (reference lidm/modules/diffusion/openaimodel.py:445-470)."""
    image_size: Tuple[int, int] = (16, 128)
    in_channels: int = 8
    out_channels: int = 8
    model_channels: int = 256
    attention_resolutions: Tuple[int, ...] = (4, 2, 1)
    num_res_blocks: int = 2
    channel_mult: Tuple[int, ...] = (1, 2, 4)
    num_head_channels: int = 32
    num_heads: int = -1
    lib_name: str = "lidm"
    use_spatial_transformer: bool = False
    transformer_depth: int = 1
    context_dim: Optional[int] = None
    use_scale_shift_norm: bool = False
    resblock_updown: bool = False
    conv_resample: bool = True
    dropout: float = 0.0
    # "openai" = lidm.modules.diffusion.openaimodel.UNetModel; "layout" = LayoutDiffusionUNetModel
    # (lidm/modules/unets/object_cross_unet.py:632-952): attention_resolutions holds its attention_ds, encoder_channels the
    # layout encoder's hidden width, use_scale_shift_norm / resblock_updown are True as shipped
    unet_type: str = "openai"
    encoder_channels: int = 0
    num_attention_blocks: int = 1
    # "efficient" = lidm.modules.unets.efficient_unet.EfficientUNet (the R2DM pixel-space denoiser): image_size = resolution,
    # model_channels = base_channels, channel_mult = channel_multiplier, num_heads = attn_num_heads, plus:
    num_residual_blocks: Tuple[int, ...] = (3, 3, 3, 3)
    gn_num_groups: int = 8
    gn_eps: float = 1e-6

    @property
    def time_embed_dim(self) -> int:
        return self.model_channels * 4


@dataclass
class AEConfig:
    """Mirror of VQModelInterface/Decoder arguments (reference lidm/models/ae/autoencoder.py:15-68,
    lidm/modules/diffusion/model_lidm.py:315-383)."""
    embed_dim: int = 8
    n_embed: int = 16384
    use_mask: bool = False
    z_channels: int = 8
    in_channels: int = 1
    out_ch: int = 1
    ch: int = 64
    ch_mult: Tuple[int, ...] = (1, 2, 2, 4)
    strides: Tuple[Tuple[int, int], ...] = ((1, 2), (2, 2), (2, 2))
    num_res_blocks: int = 2
    attn_levels: Tuple[int, ...] = ()
    lib_name: str = "lidm"


@dataclass
class LayoutEncoderConfig:
    """cond_stage_config of the layout-conditioned LiDM: LayoutTransformerEncoder.__init__
    (reference lidm/modules/encoders/layout_encoder.py:140-220), shipped condition types only."""
    layout_length: int = 13
    hidden_dim: int = 256
    output_dim: int = 1024
    num_layers: int = 6
    num_heads: int = 8
    use_final_ln: bool = True
    num_classes_for_layout_object: int = 9
    feature_map_size: Tuple[int, int] = (8, 128)
    resolution_to_attention: Tuple[int, ...] = (4, 2, 1)
    used_condition_types: Tuple[str, ...] = ("obj_class", "obj_bbox", "is_valid_obj")
    use_positional_embedding: bool = False
    use_key_padding_mask: bool = False
    not_use_layout_fusion_module: bool = False


@dataclass
class DatasetConfig:
    """`data.params.dataset` block (what range2pcd / range2xyz are called with, scripts/sample.py:29-35)."""
    size: Tuple[int, int] = (64, 1024)
    fov: Tuple[float, float] = (3.0, -25.0)
    depth_range: Tuple[float, float] = (1.0, 56.0)
    depth_scale: float = 5.84
    log_scale: bool = True


@dataclass
class LidmConfig:
    timesteps: int = 1000
    linear_start: float = 0.0015
    linear_end: float = 0.0195
    beta_schedule: str = "linear"
    channels: int = 8
    image_size: Tuple[int, int] = (16, 128)
    scale_factor: float = 1.0
    parameterization: str = "eps"
    conditioning_key: Optional[str] = None
    # numeric mode of the CUDA path (not a reference option), U-Net: "bf16" = plain bf16 tensor-core GEMMs (north_star
    # bf16 budget, eps within 2e-2); "fp32" = precise mode, every GEMM as a 3-way bf16 operand split with an fp32
    # residual stream (north_star fp32 bars: eps within 1e-3, final image within 1e-2), about 3x the GEMM work;
    # "fp16" = IEEE-half operands and activations at the bf16 tensor rate (saturating conversions)
    precision: str = "bf16"
    # numeric mode of the first stage (decoder / encoder): "bf16" | "fp32" | "fp16"; None = "fp16" under a bf16 U-Net (the
    # decoder's bf16 rounding alone is 2.4e-2 on the final image, above north_star's 1e-2; half precision gives 3e-3 at
    # the same speed), otherwise the U-Net's mode
    ae_precision: Optional[str] = None

    @property
    def ae_precision_resolved(self) -> str:
        if self.ae_precision is not None:
            return self.ae_precision
        return "fp16" if self.precision == "bf16" else self.precision
    unet: UNetConfig = field(default_factory=UNetConfig)
    ae: AEConfig = field(default_factory=AEConfig)
    dataset: DatasetConfig = field(default_factory=DatasetConfig)
    layout_encoder: Optional[LayoutEncoderConfig] = None     # cond stage of layout_crossattn models

    @property
    def latent_shape(self) -> Tuple[int, int, int]:
        return (self.channels, self.image_size[0], self.image_size[1])


def _tup(x):
    if isinstance(x, (list, tuple)):
        return tuple(_tup(v) for v in x)
    return x


def _pick(cls, params: dict):
    names = {f.name for f in dataclasses.fields(cls)}
    return {k: _tup(v) for k, v in params.items() if k in names}


def from_reference_dict(cfg: dict) -> LidmConfig:
    """Build a LidmConfig from a parsed reference YAML (dict with `model:` and optional `data:`)."""
    model = cfg["model"]
    target = model.get("target", "")
    if not (target.endswith("LatentDiffusion") or target.endswith("R2DMDiffusion")):
        raise ValueError(f"unsupported model target {target!r}: only LatentDiffusion / R2DMDiffusion are on the B200 path")
    p = model["params"]
    unet_p = p["unet_config"]["params"]
    utarget = p["unet_config"]["target"]
    if target.endswith("R2DMDiffusion"):
        return _r2dm_from_reference_dict(cfg)
    if utarget.endswith("object_cross_unet.LayoutDiffusionUNetModel"):
        if unet_p.get("attention_block_type", "GLIDE") != "ObjectAwareCrossAttention":
            raise ValueError("LayoutDiffusionUNetModel: only attention_block_type 'ObjectAwareCrossAttention' is supported")
        for k, want in (("use_positional_embedding_for_attention", True), ("use_key_padding_mask", False),
                        ("norm_first", False), ("norm_for_obj_embedding", False),
                        ("channels_scale_for_positional_embedding", 1.0)):
            if unet_p.get(k, want) != want:
                raise ValueError(f"LayoutDiffusionUNetModel: {k}={unet_p[k]!r} is not supported (shipped value: {want!r})")
        kw_u = _pick(UNetConfig, unet_p)
        kw_u.update(unet_type="layout", attention_resolutions=_tup(unet_p["attention_ds"]), lib_name="ldm")
        unet = UNetConfig(**kw_u)
    elif utarget.endswith("openaimodel.UNetModel"):
        unet = UNetConfig(**_pick(UNetConfig, unet_p))
    else:
        raise ValueError("unsupported unet target " + utarget)
    fs = p["first_stage_config"]
    if not fs["target"].endswith("VQModelInterface"):
        raise ValueError("unsupported first stage target " + fs["target"])
    fsp = fs["params"]
    ae_kw = _pick(AEConfig, fsp)
    ae_kw.update(_pick(AEConfig, fsp["ddconfig"]))
    ae = AEConfig(**ae_kw)
    cond = p.get("cond_stage_config", "__is_unconditional__")
    conditioning_key = p.get("conditioning_key", None)
    if cond == "__is_unconditional__":
        conditioning_key = None
    elif conditioning_key is None:
        conditioning_key = "concat" if p.get("concat_mode", True) else "crossattn"
    kw = dict(
        timesteps=p.get("timesteps", 1000),
        linear_start=p.get("linear_start", 1e-4),
        linear_end=p.get("linear_end", 2e-2),
        beta_schedule=p.get("beta_schedule", "linear"),
        channels=p.get("channels", 3),
        image_size=_tup(p.get("image_size", (256, 256))),
        scale_factor=p.get("scale_factor", 1.0),
        parameterization=p.get("parameterization", "eps"),
        conditioning_key=conditioning_key,
        unet=unet, ae=ae,
    )
    if isinstance(cond, dict) and str(cond.get("target", "")).endswith("layout_encoder.LayoutTransformerEncoder"):
        le = LayoutEncoderConfig(**_pick(LayoutEncoderConfig, cond["params"]))
        if set(le.used_condition_types) != {"obj_class", "obj_bbox", "is_valid_obj"} or le.use_positional_embedding \
                or le.use_key_padding_mask or le.not_use_layout_fusion_module:
            raise ValueError("LayoutTransformerEncoder: only the shipped options are supported (obj_class / obj_bbox / "
                             "is_valid_obj, no positional embedding, no key padding mask)")
        kw["layout_encoder"] = le
    ds = DatasetConfig()
    try:
        ds = DatasetConfig(**_pick(DatasetConfig, cfg["data"]["params"]["dataset"]))
    except (KeyError, TypeError):
        pass
    return LidmConfig(dataset=ds, **kw)


def _r2dm_from_reference_dict(cfg: dict) -> LidmConfig:
    """R2DMDiffusion (lidm/models/diffusion/ddpm_r2dm.py) + EfficientUNet (efficient_unet.py:188-260): pixel space, no
    first stage."""
    p = cfg["model"]["params"]
    up = p["unet_config"]["params"]
    if not p["unet_config"]["target"].endswith("efficient_unet.EfficientUNet"):
        raise ValueError("unsupported unet target " + p["unet_config"]["target"])
    if up.get("coords_encoding", "spherical_harmonics") != "fourier_features" or not up.get("ring", True):
        raise ValueError("EfficientUNet: only coords_encoding 'fourier_features' with ring padding is supported (as shipped)")
    if up.get("temb_channels") not in (None, 4 * up.get("base_channels", 128)) or up.get("out_channels") not in (None, up["in_channels"]):
        raise ValueError("EfficientUNet: temb_channels / out_channels must keep their defaults")
    mult = up.get("channel_multiplier", (1, 2, 4, 8))
    nres = up.get("num_residual_blocks", (3, 3, 3, 3))
    mult = tuple(mult) if isinstance(mult, (list, tuple)) else (mult,) * 4
    nres = tuple(nres) if isinstance(nres, (list, tuple)) else (nres,) * 4
    res = up["resolution"]
    res = tuple(res) if isinstance(res, (list, tuple)) else (res, res)
    unet = UNetConfig(image_size=res, in_channels=up["in_channels"], out_channels=up["in_channels"],
                      model_channels=up.get("base_channels", 128), channel_mult=mult, num_residual_blocks=nres,
                      gn_num_groups=up.get("gn_num_groups", 8), gn_eps=float(up.get("gn_eps", 1e-6)),
                      num_heads=up.get("attn_num_heads", 8), unet_type="efficient", attention_resolutions=(), lib_name="r2dm")
    if p.get("cond_stage_config", "__is_unconditional__") != "__is_unconditional__":
        raise ValueError("R2DMDiffusion: only the unconditional model is supported")
    kw = dict(timesteps=p.get("timesteps", 1000), linear_start=p.get("linear_start", 1e-4), linear_end=p.get("linear_end", 2e-2),
              beta_schedule=p.get("beta_schedule", "linear"), channels=p.get("channels", up["in_channels"]),
              image_size=_tup(p.get("image_size", res)), scale_factor=1.0, parameterization=p.get("parameterization", "eps"),
              conditioning_key=None, unet=unet, ae=AEConfig(ch=0, ch_mult=(), strides=()), precision="fp16")
    ds = DatasetConfig()
    try:
        ds = DatasetConfig(**_pick(DatasetConfig, cfg["data"]["params"]["dataset"]))
    except (KeyError, TypeError):
        pass
    return LidmConfig(dataset=ds, **kw)


def nuscenes_r2dm(resolution=(32, 1024)) -> LidmConfig:
    """The R2DM pixel-space model (reference models/lidm/nuscenes/r2dm/config.yaml): EfficientUNet on 2-channel (depth,
    reflectance) range images, 1024 diffusion steps; BASELINE config 5 also runs it at 64x1024."""
    H, W = resolution
    # default numeric mode fp16: the pixel-space net is deep (24 residual blocks) and its output IS the image - bf16 sits at
    # 1.9e-2 per-step eps error on the shipped size (inside north_star's 2e-2, without margin), IEEE half at 2.6e-3
    return LidmConfig(timesteps=1024, linear_start=0.0015, linear_end=0.0195, channels=2, image_size=(H, W), precision="fp16",
                      unet=UNetConfig(image_size=(H, W), in_channels=2, out_channels=2, model_channels=64,
                                      channel_mult=(1, 2, 4, 8), num_residual_blocks=(3, 3, 3, 3), gn_num_groups=8, gn_eps=1e-6,
                                      num_heads=8, unet_type="efficient", attention_resolutions=(), lib_name="r2dm"),
                      ae=AEConfig(ch=0, ch_mult=(), strides=()),
                      dataset=DatasetConfig(size=(H, W), fov=(10.0, -30.0)))


def tiny_r2dm() -> LidmConfig:
    """R2DM at 16x512 with one residual block per level (same four levels, both attention head widths): for fast tests."""
    c = nuscenes_r2dm((16, 512))
    return dataclasses.replace(c, unet=dataclasses.replace(c.unet, num_residual_blocks=(1, 1, 1, 1)))


def from_yaml(path: str) -> LidmConfig:
    import yaml
    with open(path) as f:
        return from_reference_dict(yaml.safe_load(f))


def kitti_cam2lidar() -> LidmConfig:
    """Cross-attention conditioned KITTI-360 LiDM (reference models/lidm/kitti/{cam2lidar,text2lidar}/config.yaml):
    same U-Net with SpatialTransformer blocks in place of AttentionBlocks, context_dim 512."""
    return LidmConfig(conditioning_key="crossattn",
                      unet=UNetConfig(use_spatial_transformer=True, context_dim=512),
                      dataset=DatasetConfig(depth_scale=56.0, log_scale=False))


def kitti_sem2lidar() -> LidmConfig:
    """Concat-conditioned KITTI-360 LiDM (reference models/lidm/kitti/sem2lidar/config.yaml): the 8-channel
    rescaled semantic map is concatenated to the latent, U-Net in_channels 16."""
    return LidmConfig(conditioning_key="concat", unet=UNetConfig(in_channels=16))


def nuscenes_layout2lidar() -> LidmConfig:
    """Layout-conditioned nuScenes LiDM (reference models/lidm/nuscenes/layout2lidar/config.yaml): 32-beam range images
    (32x1024 -> 8x128 latents), LayoutDiffusionUNetModel with ObjectAwareCrossAttention at ds 2 and 4, 13 layout tokens."""
    return LidmConfig(conditioning_key="layout_crossattn", linear_end=0.0205, image_size=(8, 128),
                      unet=UNetConfig(image_size=(8, 128), unet_type="layout", model_channels=256, encoder_channels=256,
                                      num_head_channels=64, attention_resolutions=(8, 4, 2), channel_mult=(1, 2, 4),
                                      num_res_blocks=2, use_scale_shift_norm=True, resblock_updown=True, lib_name="ldm"),
                      layout_encoder=LayoutEncoderConfig(),
                      dataset=DatasetConfig(size=(32, 1024), fov=(10.0, -30.0)))


def tiny_layout() -> LidmConfig:
    """The layout2lidar structure at a small width (same three resolutions 8x128 / 4x64 / 2x32, attention at ds 2 and 4,
    13 layout tokens) with the small first stage of `tiny`: for fast tests."""
    return LidmConfig(conditioning_key="layout_crossattn", linear_end=0.0205, image_size=(8, 128),
                      unet=UNetConfig(image_size=(8, 128), unet_type="layout", model_channels=64, encoder_channels=64,
                                      num_head_channels=64, attention_resolutions=(4, 2), channel_mult=(1, 2, 4),
                                      num_res_blocks=1, use_scale_shift_norm=True, resblock_updown=True, lib_name="ldm"),
                      ae=AEConfig(n_embed=512, ch=64, ch_mult=(1, 2, 2), strides=((1, 2), (2, 2)), num_res_blocks=1),
                      layout_encoder=LayoutEncoderConfig(hidden_dim=64, output_dim=256, num_layers=2, num_heads=4),
                      dataset=DatasetConfig(size=(16, 512), fov=(10.0, -30.0)))


def kitti_uncond() -> LidmConfig:
    """The released unconditional KITTI-360 LiDM config (reference models/lidm/kitti/uncond/config.yaml)."""
    return LidmConfig()


def tiny(image_size=(8, 64), model_channels=64, channel_mult=(1, 2), n_embed=512, cond: Optional[str] = None,
         context_dim: int = 64, concat_channels: int = 4) -> LidmConfig:
    """A small same-topology config for fast tests (same op types, fewer channels/levels).
    cond: None (unconditional), "crossattn" (SpatialTransformer blocks) or "concat"."""
    kw = {}
    if cond == "crossattn":
        kw = dict(use_spatial_transformer=True, context_dim=context_dim)
    elif cond == "concat":
        kw = dict(in_channels=8 + concat_channels)
    elif cond is not None:
        raise ValueError(cond)
    unet = UNetConfig(image_size=image_size, model_channels=model_channels, channel_mult=channel_mult,
                      attention_resolutions=(2, 1), num_res_blocks=1, num_head_channels=32, **kw)
    ae = AEConfig(n_embed=n_embed, ch=64, ch_mult=(1, 2, 2), strides=((1, 2), (2, 2)), num_res_blocks=1)
    H, W = image_size
    return LidmConfig(timesteps=1000, image_size=image_size, unet=unet, ae=ae, conditioning_key=cond,
                      dataset=DatasetConfig(size=(H * 2, W * 4)))
