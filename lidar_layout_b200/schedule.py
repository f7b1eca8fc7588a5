"""Diffusion / DDIM schedules (host side, tiny).  Same arithmetic and dtypes as the reference:
make_beta_schedule / make_ddim_timesteps / make_ddim_sampling_parameters (lidm/modules/basic.py:147-197),
DDPM.register_schedule (lidm/models/diffusion/ddpm.py:120-160), DDIMSampler.make_schedule
(lidm/models/diffusion/ddim.py:26-55) and the float32 rounding of torch.full in p_sample_ddim (:191-194)."""
from __future__ import annotations

import numpy as np
import torch


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    if schedule == "linear":
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    elif schedule == "cosine":
        timesteps = torch.arange(n_timestep + 1, dtype=torch.float64) / n_timestep + cosine_s
        alphas = timesteps / (1 + cosine_s) * np.pi / 2
        alphas = torch.cos(alphas).pow(2)
        alphas = alphas / alphas[0]
        betas = 1 - alphas[1:] / alphas[:-1]
        betas = np.clip(betas, a_min=0, a_max=0.999)
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas.numpy()


def ddpm_buffers(beta_schedule, timesteps, linear_start, linear_end, cosine_s=8e-3, v_posterior=0.0):
    """float32 buffers DDPM registers; keys as in the reference state_dict."""
    betas = make_beta_schedule(beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end,
                               cosine_s=cosine_s)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    t = lambda a: torch.tensor(a, dtype=torch.float32)
    post_var = (1 - v_posterior) * betas * (1.0 - ac_prev) / (1.0 - ac) + v_posterior * betas
    return dict(
        betas=t(betas), alphas_cumprod=t(ac), alphas_cumprod_prev=t(ac_prev),
        sqrt_alphas_cumprod=t(np.sqrt(ac)), sqrt_one_minus_alphas_cumprod=t(np.sqrt(1.0 - ac)),
        log_one_minus_alphas_cumprod=t(np.log(1.0 - ac)),
        sqrt_recip_alphas_cumprod=t(np.sqrt(1.0 / ac)), sqrt_recipm1_alphas_cumprod=t(np.sqrt(1.0 / ac - 1)),
        posterior_variance=t(post_var),
        posterior_log_variance_clipped=t(np.log(np.maximum(post_var, 1e-20))),
        posterior_mean_coef1=t(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        posterior_mean_coef2=t((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
    )


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=False):
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        ddim_timesteps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * .8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    return ddim_timesteps + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=False):
    """alphacums: float32 CPU tensor.  Returns (sigmas [torch f64], alphas [torch f32], alphas_prev [np f64]) exactly
    like the reference (an out-of-range timestep raises IndexError there too, e.g. S=3 on T=1000)."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def ddim_table(alphas_cumprod_f32_cpu: torch.Tensor, ddim_timesteps, eta):
    """(n,4) float32 [a_t, a_prev, sigma_t, sqrt_one_minus_at], each rounded like torch.full((b,1,1,1), v)."""
    sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(alphas_cumprod_f32_cpu, ddim_timesteps, eta)
    sqrt_1m = np.sqrt(1.0 - alphas)
    n = len(ddim_timesteps)
    table = np.zeros((n, 4), dtype=np.float32)
    for i in range(n):
        table[i, 0] = torch.full((1,), alphas[i]).item()
        table[i, 1] = torch.full((1,), float(alphas_prev[i])).item()
        table[i, 2] = torch.full((1,), float(sigmas[i])).item()
        table[i, 3] = torch.full((1,), sqrt_1m[i]).item()
    return table, (sigmas, alphas, alphas_prev, sqrt_1m)
