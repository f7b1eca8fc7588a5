"""Python owner of one `lidm_handle` (C ABI, include/lidm_b200.h): config marshalling, state-dict feeding and the
raw tensor-in/tensor-out calls.  PyTorch is used only for device memory and the current CUDA stream."""
from __future__ import annotations

import ctypes
from ctypes import c_int32, c_int64, c_void_p
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from .config import LidmConfig


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def to_cconfig(cfg: LidmConfig) -> _lib.CConfig:
    u, a = cfg.unet, cfg.ae
    if cfg.conditioning_key not in (None, "concat", "crossattn", "hybrid", "layout_crossattn"):
        raise NotImplementedError(f"conditioning_key {cfg.conditioning_key!r} is not on the B200 path "
                                  "(supported: None, 'concat', 'crossattn', 'hybrid', 'layout_crossattn')")
    if u.unet_type == "efficient":
        return _eff_cconfig(cfg)
    layout = u.unet_type == "layout"
    if layout != (cfg.conditioning_key == "layout_crossattn"):
        raise ValueError("conditioning_key 'layout_crossattn' goes with LayoutDiffusionUNetModel (and vice versa)")
    if layout:
        if not (u.use_scale_shift_norm and u.resblock_updown):
            raise NotImplementedError("LayoutDiffusionUNetModel: use_scale_shift_norm and resblock_updown must be True (as shipped)")
        if u.encoder_channels <= 0:
            raise ValueError("LayoutDiffusionUNetModel needs encoder_channels")
    elif u.use_scale_shift_norm or u.resblock_updown or not u.conv_resample:
        raise NotImplementedError("unsupported UNetModel option for the B200 path")
    if tuple(u.image_size) != tuple(cfg.image_size) or u.in_channels < cfg.channels:
        raise ValueError("unet image_size/in_channels must match the latent shape")
    if (u.in_channels > cfg.channels) != (cfg.conditioning_key in ("concat", "hybrid")):
        raise ValueError("unet in_channels exceeds the latent channels only for concat / hybrid conditioning")
    if u.use_spatial_transformer and (u.context_dim is None or isinstance(u.context_dim, (list, tuple))):
        raise NotImplementedError("use_spatial_transformer needs a single integer context_dim")
    if bool(u.use_spatial_transformer) != (cfg.conditioning_key in ("crossattn", "hybrid")):
        raise ValueError("crossattn / hybrid conditioning needs use_spatial_transformer (and vice versa)")
    c = _lib.CConfig()
    c.in_channels, c.out_channels, c.model_channels = u.in_channels, u.out_channels, u.model_channels
    c.num_res_blocks, c.num_head_channels = u.num_res_blocks, u.num_head_channels
    c.n_channel_mult = len(u.channel_mult)
    for i, m in enumerate(u.channel_mult):
        c.channel_mult[i] = m
    c.n_attention_resolutions = len(u.attention_resolutions)
    for i, m in enumerate(u.attention_resolutions):
        c.attention_resolutions[i] = m
    c.latent_h, c.latent_w = cfg.image_size
    c.embed_dim, c.n_embed, c.z_channels = a.embed_dim, a.n_embed, a.z_channels
    c.ae_ch, c.ae_out_ch, c.ae_num_res_blocks, c.ae_use_mask = a.ch, a.out_ch, a.num_res_blocks, int(a.use_mask)
    c.ae_n_ch_mult = len(a.ch_mult)
    for i, m in enumerate(a.ch_mult):
        c.ae_ch_mult[i] = m
    for i, s in enumerate(a.strides):
        c.ae_strides[i][0], c.ae_strides[i][1] = s
    c.scale_factor = float(cfg.scale_factor)
    modes = {"bf16": 0, "fp32": 1, "fp16": 2}     # LIDM_PREC_BF16 / BF16X3 / FP16
    if cfg.precision not in modes or cfg.ae_precision_resolved not in modes:
        raise ValueError("precision / ae_precision must be one of 'bf16', 'fp32', 'fp16'")
    c.precision = modes[cfg.precision]
    c.ae_precision = modes[cfg.ae_precision_resolved] + 1
    c.latent_channels = cfg.channels
    c.use_spatial_transformer = int(bool(u.use_spatial_transformer))
    c.context_dim = int(u.context_dim or 0)
    c.transformer_depth = int(u.transformer_depth)
    c.ae_in_channels = int(a.in_channels)
    c.unet_type = 1 if layout else 0
    c.encoder_channels = int(u.encoder_channels)
    c.num_attention_blocks = int(u.num_attention_blocks)
    le = cfg.layout_encoder
    if layout and le is not None:
        if le.hidden_dim != u.encoder_channels or le.output_dim != u.time_embed_dim:
            raise ValueError("layout encoder: hidden_dim must equal the U-Net's encoder_channels, output_dim 4*model_channels")
        c.enc_layers, c.enc_heads, c.enc_out_dim = int(le.num_layers), int(le.num_heads), int(le.output_dim)
        c.enc_num_classes = int(le.num_classes_for_layout_object)
    return c


def _eff_cconfig(cfg: LidmConfig) -> _lib.CConfig:
    """struct lidm_config for the pixel-space R2DM model (unet_type 2): no first stage."""
    u = cfg.unet
    if cfg.conditioning_key is not None:
        raise NotImplementedError("R2DM: only the unconditional model is on the B200 path")
    if tuple(u.image_size) != tuple(cfg.image_size) or u.in_channels != cfg.channels or len(u.channel_mult) != 4:
        raise ValueError("R2DM: unet resolution / channels must match the image, four levels")
    modes = {"bf16": 0, "fp16": 2}
    if cfg.precision not in modes:
        raise ValueError("R2DM: precision must be 'bf16' or 'fp16'")
    c = _lib.CConfig()
    c.unet_type = 2
    c.in_channels = c.out_channels = c.latent_channels = u.in_channels
    c.model_channels = u.model_channels
    c.n_channel_mult = 4
    for i, m in enumerate(u.channel_mult):
        c.channel_mult[i] = m
    for i, n in enumerate(u.num_residual_blocks):
        c.eff_res_blocks[i] = n
    c.eff_gn_groups, c.eff_attn_heads, c.eff_gn_eps = int(u.gn_num_groups), int(u.num_heads), float(u.gn_eps)
    c.latent_h, c.latent_w = cfg.image_size
    c.scale_factor = 1.0
    c.precision = modes[cfg.precision]
    c.ae_precision = 0
    return c


class Engine:
    """One packed model on one CUDA device."""

    def __init__(self, cfg: LidmConfig, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise _lib.LidmError(-2, "CUDA is not available: lidar_layout_b200 has no CPU fallback")
        self.cfg = cfg
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._lib = _lib.load()
        self._h = c_void_p()
        with torch.cuda.device(self.device):
            cc = to_cconfig(cfg)
            _lib.check(self._lib.lidm_create(ctypes.byref(cc), ctypes.byref(self._h)))
        self.finalized = False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.lidm_destroy(h)
            self._h = c_void_p()

    # ---- weights --------------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, "torch.Tensor"], use_ema: bool = False):
        """Feed every relevant tensor of a reference state_dict, then pack (load_state_dict + ema_scope)."""
        with torch.cuda.device(self.device):
            for name, t in sd.items():
                if not (name.startswith("model.diffusion_model.") or name.startswith("model_ema.")
                        or name.startswith("first_stage_model.") or name.startswith("cond_stage_model.")):
                    continue
                if name.startswith("first_stage_model.loss."):
                    continue
                if isinstance(t, np.ndarray):
                    t = torch.from_numpy(t)
                if not torch.is_floating_point(t):
                    continue   # e.g. model_ema.num_updates
                t = t.detach().to(torch.float32).contiguous()
                shape = (c_int64 * max(t.dim(), 1))(*t.shape)
                _lib.check(self._lib.lidm_load_weight(self._h, name.encode(), c_void_p(t.data_ptr()), t.dim(), shape),
                           self._h)
            _lib.check(self._lib.lidm_finalize_weights(self._h, int(use_ema)), self._h)
        self.finalized = True
        return self

    # ---- raw calls ------------------------------------------------------------------------------------
    def _cond_ptrs(self, B, c_concat, context, what="conditioning"):
        """Validate / normalise the two conditioning tensors -> (c_concat, context, ctx_len); tensors are returned so the
        caller keeps them alive for the duration of the call."""
        u = self.cfg.unet
        if c_concat is not None:
            c_concat = _f32c(c_concat, "c_concat")
            exp = (B, u.in_channels - self.cfg.channels) + tuple(self.cfg.image_size)
            if tuple(c_concat.shape) != exp:
                raise ValueError(f"{what}: c_concat must have shape {exp}, got {tuple(c_concat.shape)}")
        L = 0
        if context is not None:
            context = _f32c(context, "context")
            if context.dim() != 3 or context.shape[0] != B or context.shape[2] != u.context_dim:
                raise ValueError(f"{what}: context must be (B={B}, L, {u.context_dim}), got {tuple(context.shape)}")
            L = int(context.shape[1])
        return c_concat, context, L

    # ---- layout conditioning (LayoutDiffusionUNetModel) -------------------------------------------------
    def layout_encode(self, layout: torch.Tensor) -> Dict[str, "torch.Tensor"]:
        """LayoutTransformerEncoder.forward (layout_encoder.py:222-281): layout (B, L, 13) -> the conditioning dict."""
        le = self.cfg.layout_encoder
        if le is None:
            raise ValueError("this model has no layout encoder configuration")
        layout = _f32c(layout, "layout")
        if layout.dim() != 3 or layout.shape[2] != 13:
            raise ValueError(f"layout must be (B, L, 13) = [bbox 8 | bbox_2d 4 | class 1], got {tuple(layout.shape)}")
        B, Lt, _ = layout.shape
        E, dev = le.hidden_dim, layout.device
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        out = {"xf_proj": f(B, le.output_dim), "xf_out": f(B, E, Lt), "obj_class_embedding": f(B, E, Lt),
               "obj_bbox_embedding": f(B, E, Lt)}
        rows = [int(r) for r in le.resolution_to_attention]
        fh, fw = le.feature_map_size
        tables = [f(1, E, r * int(fw / (fh / r))) for r in rows]
        n = len(rows)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lidm_layout_encode(
                self._h, layout.data_ptr(), B, Lt, out["xf_proj"].data_ptr(), out["xf_out"].data_ptr(),
                out["obj_class_embedding"].data_ptr(), out["obj_bbox_embedding"].data_ptr(), n, (c_int32 * max(n, 1))(*rows),
                (c_void_p * max(n, 1))(*[t.data_ptr() for t in tables]), _stream_ptr(self.device)), self._h)
        for r, t in zip(rows, tables):       # the reference repeat_interleaves the table over the batch: an expanded view
            out[f"image_patch_bbox_embedding_for_resolution{r}"] = t.expand(B, -1, -1)
        # is_valid_obj (layout_encoder.py:224,268): the reference keeps the class column's trailing dimension
        out["key_padding_mask"] = (1 - (layout[..., 12:13] > 0).int()).bool()
        return out

    def set_layout_cond(self, cond: Dict[str, "torch.Tensor"]):
        """Hand the output dict of LayoutTransformerEncoder.forward (layout_encoder.py:222-281) to the engine: everything
        that depends on the conditioning only is computed once here (lidm_layout_set_cond).  Cached on the identity and
        version of the dict's tensors, so a sampler that passes the same dict every step pays once."""
        need = ("xf_proj", "xf_out", "obj_class_embedding", "obj_bbox_embedding")
        for k in need:
            if k not in cond:
                raise KeyError(f"layout conditioning is missing {k!r}")
        key = tuple((k, v.data_ptr(), v._version, tuple(v.shape)) for k, v in sorted(cond.items())
                    if isinstance(v, torch.Tensor))
        if getattr(self, "_layout_key", None) == key:
            return
        t = {k: _f32c(cond[k], k) for k in need}
        B, E, Lt = t["xf_out"].shape
        if E != self.cfg.unet.encoder_channels or t["xf_proj"].shape != (B, self.cfg.unet.time_embed_dim):
            raise ValueError("layout conditioning: xf_out must be (B, encoder_channels, L), xf_proj (B, 4*model_channels)")
        rows, embs, batches, keep = [], [], [], []
        pre = "image_patch_bbox_embedding_for_resolution"
        for k, v in cond.items():
            if not k.startswith(pre):
                continue
            if v.dim() == 3 and v.stride(0) == 0:          # an expanded (batch-broadcast) table: take its single row
                v = v[:1]
            v = _f32c(v, k)
            # the reference encoder repeat_interleaves one (1, E, L1) tensor over the batch (layout_encoder.py:251-257):
            # detect that once and let the engine compute the positional projection a single time
            same = bool(v.shape[0] > 1 and torch.equal(v[:1].expand_as(v), v))
            vv = v[:1].contiguous() if same or v.shape[0] == 1 else v
            if vv.shape[0] not in (1, B):
                raise ValueError(f"{k}: batch {vv.shape[0]} does not match the conditioning batch {B}")
            rows.append(int(k[len(pre):])); embs.append(vv.data_ptr()); batches.append(int(vv.shape[0])); keep.append(vv)
        n = len(rows)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lidm_layout_set_cond(
                self._h, B, Lt, t["xf_proj"].data_ptr(), t["xf_out"].data_ptr(), t["obj_class_embedding"].data_ptr(),
                t["obj_bbox_embedding"].data_ptr(), n, (c_int32 * max(n, 1))(*rows), (c_void_p * max(n, 1))(*embs),
                (c_int32 * max(n, 1))(*batches), _stream_ptr(self.device)), self._h)
        self._layout_key = key
        self._layout_refs = list(cond.values())     # keep the keyed tensors alive: their addresses cannot be recycled

    def unet_forward(self, x: torch.Tensor, t: torch.Tensor, c_concat: Optional[torch.Tensor] = None,
                     context: Optional[torch.Tensor] = None, layout_cond: Optional[Dict] = None) -> torch.Tensor:
        if layout_cond is not None:
            self.set_layout_cond(layout_cond)
        x = _f32c(x, "x")
        t = t.to(device=x.device, dtype=torch.int64).contiguous()
        out = torch.empty_like(x)
        c_concat, context, L = self._cond_ptrs(x.shape[0], c_concat, context)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lidm_unet_forward_cond(
                self._h, x.data_ptr(), t.data_ptr(), c_concat.data_ptr() if c_concat is not None else None,
                context.data_ptr() if context is not None else None, L, out.data_ptr(), x.shape[0],
                _stream_ptr(self.device)), self._h)
        return out

    def cfg_combine(self, eps2: torch.Tensor, scale: float) -> torch.Tensor:
        """e_u + scale * (e_c - e_u) for eps2 = cat([e_u, e_c]) (ddim.py:179-180)."""
        eps2 = _f32c(eps2, "eps2")
        out = torch.empty((eps2.shape[0] // 2,) + tuple(eps2.shape[1:]), dtype=torch.float32, device=eps2.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lidm_cfg_combine(eps2.data_ptr(), float(scale), out.data_ptr(), out.numel(),
                                                  _stream_ptr(self.device)))
        return out

    def ddim_sample(self, x_T: torch.Tensor, timesteps: np.ndarray, table: np.ndarray,
                    noise: Optional[torch.Tensor] = None, temperature: float = 1.0, want_pred_x0: bool = False,
                    c_concat: Optional[torch.Tensor] = None, context: Optional[torch.Tensor] = None,
                    uncond_concat: Optional[torch.Tensor] = None, uncond_context: Optional[torch.Tensor] = None,
                    guidance_scale: float = 1.0, layout_cond: Optional[Dict] = None):
        """Whole DDIM loop on the device.  timesteps ascending int64 [n]; table float32 [n,4].  Optional conditioning
        (concat tensor and / or cross-attention context) and classifier-free guidance against its unconditional twin."""
        if layout_cond is not None:
            self.set_layout_cond(layout_cond)
        x = _f32c(x_T, "x_T").clone()
        n = int(len(timesteps))
        ts = np.ascontiguousarray(timesteps, dtype=np.int64)
        tab = np.ascontiguousarray(table, dtype=np.float32)
        assert tab.shape == (n, 4)
        nz_ptr = None
        if noise is not None:
            noise = _f32c(noise, "noise")
            assert noise.shape == (n,) + tuple(x.shape), "noise must be (n_steps, B, C, H, W)"
            nz_ptr = noise.data_ptr()
        pred = torch.empty_like(x) if want_pred_x0 else None
        B = x.shape[0]
        c_concat, context, L = self._cond_ptrs(B, c_concat, context)
        uncond_concat, uncond_context, Lu = self._cond_ptrs(B, uncond_concat, uncond_context, "unconditional conditioning")
        if uncond_context is not None and Lu != L:
            raise ValueError("the unconditional context must have the same length as the context")
        ptr = lambda t: t.data_ptr() if t is not None else None
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lidm_ddim_sample_cond(
                self._h, x.data_ptr(), ts.ctypes.data_as(ctypes.POINTER(c_int64)),
                tab.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n, nz_ptr, float(temperature),
                ptr(pred), B, ptr(c_concat), ptr(context), L, ptr(uncond_concat), ptr(uncond_context),
                float(guidance_scale), _stream_ptr(self.device)), self._h)
        return x, pred

    def vq_quantize(self, z: torch.Tensor):
        """first_stage_model.quantize(z): (z_q, indices int32 (B*h*w,))."""
        z = _f32c(z, "z")
        zq = torch.empty_like(z)
        idx = torch.empty((z.shape[0] * z.shape[2] * z.shape[3],), dtype=torch.int32, device=z.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lidm_vq_quantize(self._h, z.data_ptr(), zq.data_ptr(), idx.data_ptr(), z.shape[0],
                                                  _stream_ptr(self.device)), self._h)
        return zq, idx

    def vq_encode(self, img: torch.Tensor) -> torch.Tensor:
        """VQModelInterface.encode: quant_conv(encoder(img)), (B, in_channels, H, W) -> (B, embed_dim, h, w)."""
        img = _f32c(img, "img")
        c, H, W = self.image_shape()
        if img.dim() != 4 or tuple(img.shape[1:]) != (self.cfg.ae.in_channels, H, W):
            raise ValueError(f"img must be (B, {self.cfg.ae.in_channels}, {H}, {W}), got {tuple(img.shape)}")
        z = torch.empty((img.shape[0], self.cfg.ae.embed_dim) + tuple(self.cfg.image_size), dtype=torch.float32,
                        device=img.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lidm_vq_encode(self._h, img.data_ptr(), z.data_ptr(), img.shape[0],
                                                _stream_ptr(self.device)), self._h)
        return z

    def image_shape(self):
        c, h, w = c_int32(), c_int32(), c_int32()
        _lib.check(self._lib.lidm_image_shape(self._h, ctypes.byref(c), ctypes.byref(h), ctypes.byref(w)), self._h)
        return c.value, h.value, w.value

    def vq_decode(self, z: torch.Tensor, force_not_quantize: bool = False, return_indices: bool = False):
        z = _f32c(z, "z")
        B = z.shape[0]
        c, h, w = self.image_shape()
        img = torch.empty((B, c, h, w), dtype=torch.float32, device=z.device)
        idx = torch.empty((B * z.shape[2] * z.shape[3],), dtype=torch.int32, device=z.device) if return_indices else None
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lidm_vq_decode(self._h, z.data_ptr(), int(force_not_quantize), img.data_ptr(),
                                                idx.data_ptr() if idx is not None else None, B,
                                                _stream_ptr(self.device)), self._h)
        return (img, idx) if return_indices else img
