"""DDIMSampler with the reference's interface (lidm/models/diffusion/ddim.py:13-206), B200-native underneath.

`sample()` keeps the reference signature, return value and RNG consumption order.  The common case (no
callbacks / mask / quantize_x0 / score corrector / noise dropout) runs the whole loop on the device through one
C-ABI call per logging segment (lidm_ddim_sample: U-Net forward with the DDIM update fused into its last conv
epilogue); everything else goes step by step through model.apply_model + the lidm_ddim_step kernel.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops, schedule


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule

    def register_buffer(self, name, attr):
        if type(attr) == torch.Tensor:
            if attr.device != self.model.device:
                attr = attr.to(self.model.device)
        setattr(self, name, attr)

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=False):
        """reference ddim.py:26-55 (same buffers, same dtypes)."""
        self.ddim_timesteps = schedule.make_ddim_timesteps(ddim_discr_method=ddim_discretize,
                                                           num_ddim_timesteps=ddim_num_steps,
                                                           num_ddpm_timesteps=self.ddpm_num_timesteps, verbose=verbose)
        alphas_cumprod = self.model.alphas_cumprod
        assert alphas_cumprod.shape[0] == self.ddpm_num_timesteps, 'alphas have to be defined for each timestep'
        to_torch = lambda x: x.clone().detach().to(torch.float32).to(self.model.device)
        self.register_buffer('betas', to_torch(self.model.betas))
        self.register_buffer('alphas_cumprod', to_torch(alphas_cumprod))
        self.register_buffer('alphas_cumprod_prev', to_torch(self.model.alphas_cumprod_prev))
        ac = alphas_cumprod.cpu()
        self.register_buffer('sqrt_alphas_cumprod', to_torch(np.sqrt(ac)))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', to_torch(np.sqrt(1. - ac)))
        self.register_buffer('log_one_minus_alphas_cumprod', to_torch(np.log(1. - ac)))
        self.register_buffer('sqrt_recip_alphas_cumprod', to_torch(np.sqrt(1. / ac)))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', to_torch(np.sqrt(1. / ac - 1)))
        table, (sigmas, alphas, alphas_prev, sqrt_1m) = schedule.ddim_table(ac, self.ddim_timesteps, ddim_eta)
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = sigmas, alphas, alphas_prev
        self.ddim_sqrt_one_minus_alphas = sqrt_1m
        self.ddim_table = table            # (n,4) float32: what p_sample_ddim's torch.full would hold
        sigmas_for_original_sampling_steps = ddim_eta * torch.sqrt(
            (1 - self.alphas_cumprod_prev) / (1 - self.alphas_cumprod) * (
                    1 - self.alphas_cumprod / self.alphas_cumprod_prev))
        self.register_buffer('ddim_sigmas_for_original_num_steps', sigmas_for_original_sampling_steps)

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None,
               img_callback=None, quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0.,
               score_corrector=None, corrector_kwargs=None, verbose=False, disable_tqdm=True, x_T=None,
               log_every_t=100, unconditional_guidance_scale=1., unconditional_conditioning=None, **kwargs):
        if conditioning is not None:
            if isinstance(conditioning, dict):
                cbs = conditioning[list(conditioning.keys())[0]].shape[0]
                if cbs != batch_size:
                    print(f"Warning: Got {cbs} conditionings but batch-size is {batch_size}")
            else:
                if conditioning.shape[0] != batch_size:
                    print(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        size = (batch_size, C, H, W)
        return self.ddim_sampling(conditioning, size, callback=callback, img_callback=img_callback,
                                  quantize_denoised=quantize_x0, mask=mask, x0=x0, ddim_use_original_steps=False,
                                  noise_dropout=noise_dropout, temperature=temperature,
                                  score_corrector=score_corrector, corrector_kwargs=corrector_kwargs, x_T=x_T,
                                  log_every_t=log_every_t,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, verbose=verbose,
                                  disable_tqdm=disable_tqdm)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, ddim_use_original_steps=False, callback=None, timesteps=None,
                      quantize_denoised=False, mask=None, x0=None, img_callback=None, log_every_t=100,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, verbose=False,
                      disable_tqdm=True):
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T.to(device)
        if ddim_use_original_steps:
            raise NotImplementedError("ddim_use_original_steps is not supported on the B200 path")
        if timesteps is None:
            timesteps = self.ddim_timesteps
            table = self.ddim_table
        else:
            subset_end = int(min(timesteps / self.ddim_timesteps.shape[0], 1) * self.ddim_timesteps.shape[0]) - 1
            timesteps = self.ddim_timesteps[:subset_end]
            table = self.ddim_table[:subset_end]
        total_steps = timesteps.shape[0]
        intermediates = {'x_inter': [img], 'pred_x0': [img]}
        fused = (callback is None and img_callback is None and not quantize_denoised
                 and mask is None and score_corrector is None and noise_dropout == 0.)
        if fused:
            # conditioning tensors for the on-device loop (+ their unconditional twins for classifier-free guidance)
            cond_kw = {}
            if getattr(getattr(self.model, "model", None), "conditioning_key", None) == "layout_crossattn":
                if unconditional_conditioning is not None and unconditional_guidance_scale != 1.:
                    raise NotImplementedError("classifier-free guidance is not wired for the layout U-Net")
                cond_kw = dict(layout_cond=cond)
            elif cond is not None:
                c_concat, context = self.model.split_conditioning(cond)
                guided = unconditional_conditioning is not None and unconditional_guidance_scale != 1.
                uc_concat, uc_context = (self.model.split_conditioning(unconditional_conditioning) if guided
                                         else (None, None))
                cond_kw = dict(c_concat=c_concat, context=context, uncond_concat=uc_concat, uncond_context=uc_context,
                               guidance_scale=float(unconditional_guidance_scale) if guided else 1.0)
            # the reference draws one randn(shape) per step in loop order (ddim.py:202) whatever eta is
            noise = torch.stack([torch.randn(shape, device=device) for _ in range(total_steps)])
            use_noise = bool(np.any(table[:, 2] != 0))
            # split the loop where the reference records intermediates (ddim.py:161-163)
            log_at = [idx for idx in range(total_steps - 1, -1, -1)
                      if idx % log_every_t == 0 or idx == total_steps - 1]
            hi = total_steps - 1
            for stop in log_at:
                i0, i1 = total_steps - 1 - hi, total_steps - 1 - stop      # loop iterations [i0, i1]
                img, pred_x0 = self.model.engine.ddim_sample(
                    img, timesteps[stop:hi + 1], table[stop:hi + 1],
                    noise=noise[i0:i1 + 1] if use_noise else None, temperature=temperature, want_pred_x0=True,
                    **cond_kw)
                intermediates['x_inter'].append(img)
                intermediates['pred_x0'].append(pred_x0)
                hi = stop - 1
            return img, intermediates

        time_range = np.flip(timesteps)
        for i, step in enumerate(time_range):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            if mask is not None:
                assert x0 is not None
                img_orig = self.model.q_sample(x0, ts)
                img = img_orig * mask + (1. - mask) * img
            img, pred_x0 = self.p_sample_ddim(img, cond, ts, index=index, quantize_denoised=quantize_denoised,
                                              temperature=temperature, noise_dropout=noise_dropout,
                                              score_corrector=score_corrector, corrector_kwargs=corrector_kwargs,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning,
                                              table=table)
            if callback: callback(i)
            if img_callback: img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates['x_inter'].append(img)
                intermediates['pred_x0'].append(pred_x0)
        return img, intermediates

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, table=None):
        """reference ddim.py:167-206, one step (eps from the U-Net kernels, update from lidm_ddim_step)."""
        if use_original_steps:
            raise NotImplementedError("use_original_steps is not supported on the B200 path")
        table = self.ddim_table if table is None else table
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            e_t = self.model.apply_model(x, t, c)
        else:
            # ddim.py:175-180: one 2B evaluation on [uncond | cond], then e_u + s (e_c - e_u)
            x_in = torch.cat([x] * 2)
            t_in = torch.cat([t] * 2)
            if isinstance(c, dict):
                c_in = {k: [torch.cat([unconditional_conditioning[k][i], c[k][i]]) for i in range(len(c[k]))] for k in c}
            else:
                c_in = torch.cat([unconditional_conditioning, c])
            e_t = self.model.engine.cfg_combine(self.model.apply_model(x_in, t_in, c_in), unconditional_guidance_scale)
        if score_corrector is not None:
            assert self.model.parameterization == "eps"
            e_t = score_corrector.modify_score(self.model, e_t, x, t, c, **corrector_kwargs)
        coef = table[index]
        if repeat_noise:
            noise = torch.randn((1, *x.shape[1:]), device=x.device).repeat(x.shape[0], *((1,) * (len(x.shape) - 1)))
        else:
            noise = torch.randn(x.shape, device=x.device)
        if quantize_denoised or noise_dropout > 0.:
            # uncommon branches: same arithmetic in torch, element for element (ddim.py:196-205)
            a_t = torch.full((x.shape[0], 1, 1, 1), float(coef[0]), device=x.device)
            a_prev = torch.full((x.shape[0], 1, 1, 1), float(coef[1]), device=x.device)
            sigma_t = torch.full((x.shape[0], 1, 1, 1), float(coef[2]), device=x.device)
            sqrt_one_minus_at = torch.full((x.shape[0], 1, 1, 1), float(coef[3]), device=x.device)
            pred_x0 = (x - sqrt_one_minus_at * e_t) / a_t.sqrt()
            if quantize_denoised:
                pred_x0 = self.model.first_stage_model.quantize(pred_x0)[0]
            dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
            nz = sigma_t * noise * temperature
            if noise_dropout > 0.:
                nz = torch.nn.functional.dropout(nz, p=noise_dropout)
            return a_prev.sqrt() * pred_x0 + dir_xt + nz, pred_x0
        return ops.ddim_step(x, e_t, coef, noise, temperature)
