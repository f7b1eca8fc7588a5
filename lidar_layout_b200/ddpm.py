"""LatentDiffusion with the reference's sampling-side interface (lidm/models/diffusion/ddpm.py), B200-native.

Only what the sampling path touches is mirrored: construction from the reference YAML, load_state_dict,
ema_scope, apply_model, decode_first_stage, q_sample, sample / p_sample_loop (ancestral DDPM), sample_log, the
schedule buffers the samplers read, and the attribute chain scripts/sample.py uses
(`model.model.diffusion_model.image_size`, `.in_channels`, `model.cond_stage_model`).  Training code is out of scope.
"""
from __future__ import annotations

from contextlib import contextmanager
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import ops, schedule
from .config import LidmConfig, from_reference_dict, from_yaml
from .engine import Engine


def extract_into_tensor(a, t, x_shape):
    """lidm/modules/basic.py:219-222."""
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def noise_like(shape, device, repeat=False):
    """lidm/modules/basic.py:393-396 (the RNG draw of p_sample; module-level like the reference so callers can swap it)."""
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)


class _FirstStage:
    """Stands in for VQModelInterface on the attribute paths samplers use (`first_stage_model.quantize`)."""

    def __init__(self, owner):
        self._owner = owner

    def quantize(self, z):
        """VectorQuantizer2.forward's return convention: (z_q, loss, (perplexity, min_encodings, indices))."""
        zq, idx = self._owner.engine.vq_quantize(z)
        return zq, None, (None, None, idx.long())

    def decode(self, h, force_not_quantize=False):
        return self._owner.engine.vq_decode(h, force_not_quantize)

    def encode(self, x):
        return self._owner.engine.vq_encode(x)


class LatentDiffusion:
    """Drop-in for the sampling-side API of lidm.models.diffusion.ddpm.LatentDiffusion."""

    def __init__(self, cfg: LidmConfig, device: Optional[torch.device] = None, use_ema: bool = True,
                 precision: Optional[str] = None, ae_precision: Optional[str] = None):
        if precision is not None or ae_precision is not None:
            import dataclasses
            cfg = dataclasses.replace(cfg, precision=precision or cfg.precision,
                                      ae_precision=ae_precision or cfg.ae_precision)
        self.cfg = cfg
        self.engine = Engine(cfg, device)
        self.device = self.engine.device
        self.use_ema = use_ema
        self.parameterization = cfg.parameterization
        if self.parameterization != "eps":
            raise NotImplementedError("only eps-parameterisation is supported")
        self.channels = cfg.channels
        self.image_size = list(cfg.image_size)
        self.scale_factor = cfg.scale_factor
        self.cond_stage_model = None
        self.clip_denoised = False          # LatentDiffusion sets this (ddpm.py:466)
        self.log_every_t = 100
        self.v_posterior = 0.0
        self.first_stage_model = _FirstStage(self)
        # `model.model.diffusion_model.{image_size,in_channels}` (scripts/sample.py:91-93)
        self.model = SimpleNamespace(diffusion_model=SimpleNamespace(image_size=list(cfg.unet.image_size),
                                                                     in_channels=cfg.unet.in_channels),
                                     conditioning_key=cfg.conditioning_key)
        self.register_schedule()

    # ---- construction ---------------------------------------------------------------------------------
    @classmethod
    def from_config(cls, config, device=None, use_ema=True, precision=None, ae_precision=None) -> "LatentDiffusion":
        """config: path to a reference YAML, or the parsed dict (what OmegaConf.load would give).
        precision (U-Net) / ae_precision (first stage): None, "bf16", "fp32" (precise operand-split path) or "fp16";
        the default is a bf16 U-Net with an fp16 first stage (LidmConfig.ae_precision)."""
        cfg = from_yaml(config) if isinstance(config, str) else from_reference_dict(config)
        return cls(cfg, device, use_ema, precision, ae_precision)

    def register_schedule(self):
        """DDPM.register_schedule (ddpm.py:120-160)."""
        c = self.cfg
        bufs = schedule.ddpm_buffers(c.beta_schedule, c.timesteps, c.linear_start, c.linear_end,
                                     v_posterior=self.v_posterior)
        self.num_timesteps = int(bufs["betas"].shape[0])
        for k, v in bufs.items():
            setattr(self, k, v.to(self.device))

    def load_state_dict(self, sd, strict: bool = False):
        """model.load_state_dict(sd, strict=False) (scripts/sample.py:271).  EMA selection (ema_scope) is folded
        into the one-time weight packing: with use_ema the `model_ema.*` shadow tensors win where present."""
        if "state_dict" in sd and not any(k.startswith("model.") for k in sd):
            sd = sd["state_dict"]
        self.engine.load_state_dict(sd, use_ema=self.use_ema)
        return [], []

    def cuda(self, *a, **k):
        return self

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    @contextmanager
    def ema_scope(self, context=None):
        """DDPM.ema_scope (ddpm.py:174-187): a no-op here, the EMA weights were selected when packing."""
        yield None

    # ---- per-step hook --------------------------------------------------------------------------------
    @torch.no_grad()
    def get_learned_conditioning(self, c):
        """ddpm.py:558-569 for the layout-conditioned model: cond_stage_model(c) = LayoutTransformerEncoder.forward.
        (The CLIP / rescaler encoders of the other conditioned models stay outside the path: hand apply_model their output.)"""
        if self.model.conditioning_key != "layout_crossattn":
            raise NotImplementedError("get_learned_conditioning is on the B200 path for layout_crossattn models only")
        return self.engine.layout_encode(c.to(self.device))

    def split_conditioning(self, cond):
        """The dispatch of apply_model + DiffusionWrapper.forward (ddpm.py:900-909, 2313-2339) for the conditioning
        keys None / 'concat' / 'crossattn' / 'hybrid': cond (tensor, list or {'c_concat': [...], 'c_crossattn': [...]})
        -> (c_concat (B,Cc,H,W) or None, context (B,L,D) or None)."""
        key = self.model.conditioning_key
        if isinstance(cond, dict):
            pass
        else:
            if cond is None or (isinstance(cond, (list, tuple)) and all(c is None for c in cond)):
                cond = {}
            else:
                if not isinstance(cond, (list, tuple)):
                    cond = [cond]
                cond = {"c_concat" if key == "concat" else "c_crossattn": list(cond)}
        c_concat = cond.get("c_concat")
        c_cross = cond.get("c_crossattn")
        if key is None:
            if c_concat or c_cross:
                raise ValueError("this model is unconditional (conditioning_key None) but a conditioning was given")
            return None, None
        cc = torch.cat(list(c_concat), dim=1) if (c_concat and key in ("concat", "hybrid")) else None
        cx = torch.cat(list(c_cross), dim=1) if (c_cross and key in ("crossattn", "hybrid")) else None
        if key in ("concat", "hybrid") and cc is None:
            raise ValueError(f"conditioning_key {key!r} needs c_concat")
        if key in ("crossattn", "hybrid") and cx is None:
            raise ValueError(f"conditioning_key {key!r} needs c_crossattn")
        return cc, cx

    @torch.no_grad()
    def apply_model(self, x_noisy, t, cond, return_ids=False):
        """ddpm.py:900-1000 -> DiffusionWrapper.forward (:2313) -> UNetModel.forward."""
        assert not return_ids
        if self.model.conditioning_key == "layout_crossattn":
            # DiffusionWrapper.forward (ddpm.py:2334-2335): the whole dict of LayoutTransformerEncoder outputs goes to
            # LayoutDiffusionUNetModel.forward as layout_outputs
            if not isinstance(cond, dict) or "xf_proj" not in cond:
                raise ValueError("layout_crossattn conditioning is the dict LayoutTransformerEncoder.forward returns "
                                 "(model.get_learned_conditioning(layout) in the reference)")
            return self.engine.unet_forward(x_noisy, t, layout_cond=cond)
        c_concat, context = self.split_conditioning(cond)
        return self.engine.unet_forward(x_noisy, t, c_concat=c_concat, context=context)

    @torch.no_grad()
    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        """ddpm.py:717-775 -> VQModelInterface.decode (autoencoder.py:290-302); 1/scale_factor applied in-kernel."""
        if predict_cids:
            raise NotImplementedError("predict_cids is not used by the sampling path")
        return self.engine.vq_decode(z, force_not_quantize=force_not_quantize)

    @torch.no_grad()
    def encode_first_stage(self, x):
        """ddpm.py:837-... -> VQModelInterface.encode (autoencoder.py:285-288): quant_conv(encoder(x)), not quantised."""
        return self.engine.vq_encode(x)

    def get_first_stage_encoding(self, encoder_posterior):
        """ddpm.py:546-556 for a tensor posterior."""
        return self.scale_factor * encoder_posterior

    def q_sample(self, x_start, t, noise=None):
        """DDPM.q_sample (ddpm.py:306-309)."""
        noise = torch.randn_like(x_start) if noise is None else noise
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start +
                extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    # ---- ancestral DDPM sampling (scripts/sample.py --vanilla) ------------------------------------------
    def predict_start_from_noise(self, x_t, t, noise):
        """ddpm.py:219-223."""
        return (extract_into_tensor(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t -
                extract_into_tensor(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise)

    def q_posterior(self, x_start, x_t, t):
        """ddpm.py:225-232."""
        mean = (extract_into_tensor(self.posterior_mean_coef1, t, x_t.shape) * x_start +
                extract_into_tensor(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        var = extract_into_tensor(self.posterior_variance, t, x_t.shape)
        logvar = extract_into_tensor(self.posterior_log_variance_clipped, t, x_t.shape)
        return mean, var, logvar

    @torch.no_grad()
    def p_mean_variance(self, x, c, t, clip_denoised: bool, quantize_denoised=False, return_x0=False):
        """ddpm.py:1059-1088 (eps parameterisation)."""
        model_out = self.apply_model(x, t, c)
        x_recon = self.predict_start_from_noise(x, t=t, noise=model_out)
        if clip_denoised:
            x_recon.clamp_(-1., 1.)
        if quantize_denoised:
            x_recon, _, _ = self.first_stage_model.quantize(x_recon)
        mean, var, logvar = self.q_posterior(x_start=x_recon, x_t=x, t=t)
        return (mean, var, logvar, x_recon) if return_x0 else (mean, var, logvar)

    @torch.no_grad()
    def p_sample(self, x, c, t, clip_denoised=False, repeat_noise=False, return_x0=False, temperature=1.,
                 noise_dropout=0., quantize_denoised=False):
        """ddpm.py:1090-1119.  Without quantize_denoised the whole update after the U-Net (predict_start_from_noise,
        q_posterior, the masked noise term) is one kernel (`lidm_ddpm_step`), bit-identical to the tensor expression; the
        per-sample coefficients are gathered from the same (T,) buffers with the same torch ops as the reference."""
        b = x.shape[0]
        if quantize_denoised:
            outputs = self.p_mean_variance(x=x, c=c, t=t, clip_denoised=clip_denoised, return_x0=return_x0,
                                           quantize_denoised=quantize_denoised)
            mean, _, logvar = outputs[:3]
            noise = noise_like(x.shape, x.device, repeat_noise) * temperature
            if noise_dropout > 0.:
                noise = torch.nn.functional.dropout(noise, p=noise_dropout)
            nonzero_mask = (1 - (t == 0).float()).reshape(b, *((1,) * (len(x.shape) - 1)))
            out = mean + nonzero_mask * (0.5 * logvar).exp() * noise
            return (out, outputs[3]) if return_x0 else out
        from . import ops
        model_out = self.apply_model(x, t, c)
        noise = noise_like(x.shape, x.device, repeat_noise) * temperature
        if noise_dropout > 0.:
            noise = torch.nn.functional.dropout(noise, p=noise_dropout)
        nonzero_mask = 1 - (t == 0).float()
        coef = torch.stack([self.sqrt_recip_alphas_cumprod[t], self.sqrt_recipm1_alphas_cumprod[t],
                            self.posterior_mean_coef1[t], self.posterior_mean_coef2[t],
                            nonzero_mask * (0.5 * self.posterior_log_variance_clipped[t]).exp()], dim=1).float()
        return ops.ddpm_step(x, model_out, noise, coef, clip_denoised=clip_denoised, return_x0=return_x0)

    @torch.no_grad()
    def p_sample_loop(self, cond, shape, return_intermediates=False, x_T=None, verbose=True, callback=None,
                      timesteps=None, quantize_denoised=False, mask=None, x0=None, img_callback=None, start_T=None,
                      log_every_t=None):
        """ddpm.py:1177-1226."""
        log_every_t = log_every_t or self.log_every_t
        b = shape[0]
        img = torch.randn(shape, device=self.device) if x_T is None else x_T
        intermediates = [img]
        timesteps = self.num_timesteps if timesteps is None else timesteps
        if start_T is not None:
            timesteps = min(timesteps, start_T)
        if mask is not None:
            assert x0 is not None and x0.shape[2:3] == mask.shape[2:3]
        for i in reversed(range(0, timesteps)):
            ts = torch.full((b,), i, device=self.device, dtype=torch.long)
            img = self.p_sample(img, cond, ts, clip_denoised=self.clip_denoised, quantize_denoised=quantize_denoised)
            if mask is not None:
                img_orig = self.q_sample(x0, ts)
                img = img_orig * mask + (1. - mask) * img
            if i % log_every_t == 0 or i == timesteps - 1:
                intermediates.append(img)
            if callback: callback(i)
            if img_callback: img_callback(img, i)
        return (img, intermediates) if return_intermediates else img

    @torch.no_grad()
    def sample(self, cond, batch_size=16, return_intermediates=False, x_T=None, verbose=True, timesteps=None,
               quantize_denoised=False, mask=None, x0=None, shape=None, **kwargs):
        """ddpm.py:1228-1244."""
        if shape is None:
            shape = (batch_size, self.channels, *self.image_size)
        return self.p_sample_loop(cond, shape, return_intermediates=return_intermediates, x_T=x_T, verbose=verbose,
                                  timesteps=timesteps, quantize_denoised=quantize_denoised, mask=mask, x0=x0)

    @torch.no_grad()
    def sample_log(self, cond, batch_size, ddim, ddim_steps, **kwargs):
        """ddpm.py:1246-1258."""
        if ddim:
            from .ddim import DDIMSampler
            ddim_sampler = DDIMSampler(self)
            shape = (self.channels, *self.image_size)
            return ddim_sampler.sample(ddim_steps, batch_size, shape, cond, verbose=False, **kwargs)
        return self.sample(cond=cond, batch_size=batch_size, return_intermediates=True, **kwargs)


class R2DMDiffusion(LatentDiffusion):
    """Drop-in for the sampling-side API of lidm.models.diffusion.ddpm_r2dm.R2DMDiffusion (ddpm_r2dm.py:11-380): the
    pixel-space range-image diffusion.  Same DDPM schedule / ancestral sampler / DDIM sampler as LatentDiffusion; the
    denoiser is the EfficientUNet (efficient_unet.py:188-295) and there is no first stage."""

    def __init__(self, cfg: LidmConfig, device: Optional[torch.device] = None, use_ema: bool = True,
                 precision: Optional[str] = None):
        if cfg.unet.unet_type != "efficient":
            raise ValueError("R2DMDiffusion needs an EfficientUNet configuration")
        super().__init__(cfg, device, use_ema, precision)
        self.first_stage_model = None
        self.lidar_utils_config = dict(log_scale=cfg.dataset.log_scale, depth_range=list(cfg.dataset.depth_range))

    @classmethod
    def from_config(cls, config, device=None, use_ema=True, precision=None):
        cfg = from_yaml(config) if isinstance(config, str) else from_reference_dict(config)
        return cls(cfg, device, use_ema, precision)

    @torch.no_grad()
    def apply_model(self, x_noisy, t, cond=None, return_ids=False):
        """ddpm_r2dm.py:274-286 -> DiffusionWrapper.forward (conditioning_key None) -> EfficientUNet.forward."""
        assert not return_ids
        if cond not in (None, {}) and not (isinstance(cond, (list, tuple)) and all(c is None for c in cond)):
            raise ValueError("the R2DM model on the B200 path is unconditional")
        return self.engine.unet_forward(x_noisy, t)

    def decode_first_stage(self, z, *a, **k):
        raise NotImplementedError("R2DMDiffusion is a pixel-space model: there is no first stage")

    encode_first_stage = decode_first_stage


def range_images_to_points(x, dataset_cfg):
    """custom_to_pcd for a whole batch on the device (scripts/sample.py:29-35): x (B,1,H,W) in [-1,1] ->
    xyz (B,3,H,W) fp32 with -1 where masked, mask (B,H,W) uint8."""
    return ops.backproject(x, dataset_cfg.fov, dataset_cfg.depth_range, dataset_cfg.depth_scale,
                           dataset_cfg.log_scale, return_mask=True)
