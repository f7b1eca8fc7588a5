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
    if u.use_spatial_transformer or cfg.conditioning_key is not None:
        raise NotImplementedError("conditioned U-Nets (SpatialTransformer / concat) are not on the B200 path yet")
    if u.use_scale_shift_norm or u.resblock_updown or not u.conv_resample:
        raise NotImplementedError("unsupported UNetModel option for the B200 path")
    if tuple(u.image_size) != tuple(cfg.image_size) or u.in_channels != cfg.channels:
        raise ValueError("unet image_size/in_channels must match the latent shape")
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
    if cfg.precision not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    c.precision = 1 if cfg.precision == "fp32" else 0
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
                        or name.startswith("first_stage_model.")):
                    continue
                if name.startswith("first_stage_model.encoder.") or name.startswith("first_stage_model.loss."):
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
    def unet_forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        x = _f32c(x, "x")
        t = t.to(device=x.device, dtype=torch.int64).contiguous()
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lidm_unet_forward(self._h, x.data_ptr(), t.data_ptr(), out.data_ptr(), x.shape[0],
                                                   _stream_ptr(self.device)), self._h)
        return out

    def ddim_sample(self, x_T: torch.Tensor, timesteps: np.ndarray, table: np.ndarray,
                    noise: Optional[torch.Tensor] = None, temperature: float = 1.0, want_pred_x0: bool = False):
        """Whole DDIM loop on the device.  timesteps ascending int64 [n]; table float32 [n,4]."""
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
        with torch.cuda.device(self.device):
            _lib.check(self._lib.lidm_ddim_sample(
                self._h, x.data_ptr(), ts.ctypes.data_as(ctypes.POINTER(c_int64)),
                tab.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), n, nz_ptr, float(temperature),
                pred.data_ptr() if pred is not None else None, x.shape[0], _stream_ptr(self.device)), self._h)
        return x, pred

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
