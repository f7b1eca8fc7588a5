"""GPU (-m gpu): the reference-facing Python API (same names/arguments as the reference) end to end:
LatentDiffusion.from_config -> load_state_dict -> DDIMSampler.sample -> decode_first_stage -> range2pcd."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from lidar_layout_b200 import config as C
from lidar_layout_b200.weights import random_state_dict
from oracle import torch_ref as R
from oracle.make_golden import inputs_for, tiny_yaml


@pytest.fixture(scope="module")
def model(built_lib):
    import lidar_layout_b200 as L
    cfg = C.tiny()
    m = L.LatentDiffusion.from_config(tiny_yaml(cfg), use_ema=False)
    m.load_state_dict(random_state_dict(cfg, 0), strict=False)
    return m.cuda().eval()


def test_sampler_matches_reference_run(model, golden_tiny):
    import lidar_layout_b200 as L
    g = golden_tiny
    cfg = model.cfg
    B, S = int(g["B"]), int(g["S_short"])
    x_T, noise, z = inputs_for(cfg, B, S + 2)
    sampler = L.DDIMSampler(model)
    with model.ema_scope("Plotting"):
        samples, inter = sampler.sample(S, batch_size=B, shape=cfg.latent_shape, eta=0.0,
                                        x_T=torch.from_numpy(x_T).cuda(), verbose=False)
    assert R.rel_l2(samples.cpu(), g["ddim_eta0_final"]) < 1e-2
    assert len(inter["x_inter"]) == len(inter["pred_x0"]) and torch.equal(inter["x_inter"][-1], samples)
    # step-by-step path (callback given) must agree with the fused loop bit for bit
    seen = []
    samples2, _ = sampler.sample(S, batch_size=B, shape=cfg.latent_shape, eta=0.0, x_T=torch.from_numpy(x_T).cuda(),
                                 callback=lambda i: seen.append(i))
    assert seen == list(range(S)) and torch.equal(samples2, samples)
    img = model.decode_first_stage(samples)
    assert img.shape == (B, 1, cfg.image_size[0] * 2, cfg.image_size[1] * 4)      # strides (1,2),(2,2)


def test_rng_consumption_matches_reference_order(model):
    import lidar_layout_b200 as L
    cfg = model.cfg
    sampler = L.DDIMSampler(model)
    torch.manual_seed(1000)
    a, _ = sampler.sample(4, batch_size=2, shape=cfg.latent_shape, eta=1.0)
    # reference order (ddim.py:125,202): x_T first, then one randn(shape) per step
    torch.manual_seed(1000)
    x_T = torch.randn((2,) + tuple(cfg.latent_shape), device="cuda")
    noise = torch.stack([torch.randn((2,) + tuple(cfg.latent_shape), device="cuda") for _ in range(4)])
    b, _ = model.engine.ddim_sample(x_T, sampler.ddim_timesteps, sampler.ddim_table, noise=noise)
    assert torch.equal(a, b)


def test_range2pcd_dropin(model, golden_tiny):
    import lidar_layout_b200 as L
    g = golden_tiny
    ds = model.cfg.dataset
    unit = (np.clip(g["bp_img"], -1.0, 1.0) + 1.0) / 2.0
    pcd, color, label = L.range2pcd(unit, fov=ds.fov, depth_range=ds.depth_range, depth_scale=ds.depth_scale,
                                    log_scale=ds.log_scale)
    assert pcd.shape == g["bp_pcd"].shape and pcd.dtype == np.float64 and label is None
    assert np.abs(pcd - g["bp_pcd"]).max() < 2e-5
    xyz = L.range2xyz(unit, fov=ds.fov, depth_range=ds.depth_range, depth_scale=ds.depth_scale, log_scale=ds.log_scale)
    assert np.abs(xyz - g["bp_xyz"]).max() < 2e-5


def test_error_behaviour(model):
    import lidar_layout_b200 as L
    with pytest.raises(ValueError):     # an unconditional model handed a conditioning tensor
        model.apply_model(torch.zeros(1, 8, 8, 64).cuda(), torch.zeros(1, dtype=torch.long).cuda(), torch.zeros(1, 4, 512))
    with pytest.raises(ValueError):
        model.engine.unet_forward(torch.zeros(1, 8, 8, 64), torch.zeros(1, dtype=torch.long))     # CPU tensor
    with pytest.raises(IndexError):
        L.DDIMSampler(model).sample(3, 1, model.cfg.latent_shape)   # S=3 on T=1000 fails in the reference too


def test_ancestral_ddpm_against_reference(model, monkeypatch):
    """LatentDiffusion.sample -> p_sample_loop -> p_sample (scripts/sample.py --vanilla): the last 5
    of the 1000 steps with the reference run's per-step noise injected (tests/golden/tiny_ddpm.npz)."""
    import os
    from lidar_layout_b200 import ddpm as ddpm_mod
    from oracle import torch_ref as R
    from oracle.make_golden import inputs_for
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_ddpm.npz"))
    cfg = model.cfg
    B, steps = int(g["B"]), int(g["steps"])
    x_T, noise, _ = inputs_for(cfg, B, steps)
    it = iter([torch.from_numpy(n).cuda() for n in noise])
    monkeypatch.setattr(ddpm_mod, "noise_like", lambda shape, device, repeat=False: next(it))
    x = model.sample(None, batch_size=B, x_T=torch.from_numpy(x_T).cuda(), timesteps=steps)
    assert R.rel_l2(x.cpu(), g["final"]) < 1e-2
    # posterior update teacher-forced on the reference's (x_t, eps): the arithmetic is torch's own, bit for bit
    for i in range(steps - 1):
        t = torch.from_numpy(g["t"][i]).cuda()
        xt, e = torch.from_numpy(g["xt"][i]).cuda(), torch.from_numpy(g["eps"][i]).cuda()
        x_recon = model.predict_start_from_noise(xt, t=t, noise=e)
        mean, _, logvar = model.q_posterior(x_start=x_recon, x_t=xt, t=t)
        nz = torch.from_numpy(noise[i]).cuda()
        mask = (1 - (t == 0).float()).reshape(B, 1, 1, 1)
        xp = mean + mask * (0.5 * logvar).exp() * nz
        assert R.rel_l2(xp.cpu(), g["xt"][i + 1]) < 1e-6
        # the fused kernel behind p_sample (lidm_ddpm_step) is that expression, bit for bit (incl. clamp and the t == 0 mask)
        from lidar_layout_b200 import ops
        for tt, clip in ((t, False), (t, True), (torch.zeros_like(t), False)):
            xr = model.predict_start_from_noise(xt, t=tt, noise=e)
            if clip:
                xr = xr.clamp(-1., 1.)
            mean, _, logvar = model.q_posterior(x_start=xr, x_t=xt, t=tt)
            want = mean + (1 - (tt == 0).float()).reshape(B, 1, 1, 1) * (0.5 * logvar).exp() * nz
            coef = torch.stack([model.sqrt_recip_alphas_cumprod[tt], model.sqrt_recipm1_alphas_cumprod[tt],
                                model.posterior_mean_coef1[tt], model.posterior_mean_coef2[tt],
                                (1 - (tt == 0).float()) * (0.5 * model.posterior_log_variance_clipped[tt]).exp()], dim=1).float()
            got, got0 = ops.ddpm_step(xt, e, nz, coef, clip_denoised=clip, return_x0=True)
            assert torch.equal(got, want) and torch.equal(got0, xr)
    # and without injected noise it runs on the global RNG like the reference
    monkeypatch.undo()
    torch.manual_seed(0)
    x = model.sample(None, batch_size=2, timesteps=3)
    assert x.shape == (2,) + tuple(cfg.latent_shape) and bool(torch.isfinite(x).all())


def test_quantize_and_quantize_x0(model):
    """first_stage_model.quantize and DDIMSampler.sample(quantize_x0=True) (ddim.py:198-199) against the oracle."""
    import lidar_layout_b200 as L
    from lidar_layout_b200.weights import random_state_dict
    from oracle import torch_ref as R
    cfg = model.cfg
    sd = random_state_dict(cfg, 0)
    torch.manual_seed(3)
    z = torch.randn(2, *cfg.latent_shape)
    zq, _, (_, _, idx) = model.first_stage_model.quantize(z.cuda())
    zq_ref, idx_ref = R.vq_quantize(z, sd["first_stage_model.quantize.embedding.weight"])
    assert torch.equal(idx.cpu(), idx_ref) and torch.equal(zq.cpu(), zq_ref)        # index work: exact
    # one p_sample_ddim step with quantize_denoised, teacher-forced on the device's own eps: the update arithmetic
    # and the quantiser are exact, so x_prev must match the oracle bit for bit
    sampler = L.DDIMSampler(model)
    sampler.make_schedule(4, ddim_eta=0.0)
    x = torch.randn(2, *cfg.latent_shape)
    index = 3
    t = torch.full((2,), int(sampler.ddim_timesteps[index]), dtype=torch.long)
    e = model.apply_model(x.cuda(), t.cuda(), None)
    xp, p0 = sampler.p_sample_ddim(x.cuda(), None, t.cuda(), index=index, quantize_denoised=True)
    qf = lambda p: R.vq_quantize(p, sd["first_stage_model.quantize.embedding.weight"])[0]
    xp_ref, p0_ref = R.ddim_step(x, e.cpu(), sampler.ddim_table[index], torch.zeros_like(x), 1.0, quantize=qf)
    assert torch.equal(p0.cpu(), p0_ref) and torch.equal(xp.cpu(), xp_ref)
    # and the whole sampler accepts the option (ddim.py:198-199); the quantiser is discontinuous, so a free-running
    # comparison against the fp32 oracle is not meaningful beyond finiteness
    got, _ = sampler.sample(4, batch_size=2, shape=cfg.latent_shape, eta=0.0, x_T=x.cuda(), quantize_x0=True)
    assert torch.isfinite(got).all()


def test_ddim_inpainting_against_reference(model, monkeypatch):
    """DDIMSampler.sample(mask=, x0=) (ddim.py:146-149) with q_sample's randn_like draws injected (tiny_inpaint.npz)."""
    import os
    import lidar_layout_b200 as L
    from oracle import torch_ref as R
    from oracle.make_golden import inputs_for
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_inpaint.npz"))
    cfg = model.cfg
    B, S = int(g["B"]), int(g["S"])
    x_T, noise, x0 = inputs_for(cfg, B, S)
    it = iter([torch.from_numpy(n).cuda() for n in noise])
    monkeypatch.setattr(torch, "randn_like", lambda t, *a, **k: next(it))
    z, _ = L.DDIMSampler(model).sample(S, batch_size=B, shape=cfg.latent_shape, eta=0.0, x_T=torch.from_numpy(x_T).cuda(),
                                       mask=torch.from_numpy(g["mask"]).cuda(), x0=torch.from_numpy(x0).cuda())
    assert R.rel_l2(z.cpu(), g["final"]) < 1e-2


def test_use_ema_selects_the_shadow_weights(built_lib):
    """DDPM.ema_scope / LitEma.copy_to (ddpm.py:174-187, ema.py:46-57): with use_ema the `model_ema.*` shadow buffers (the
    parameter names with the dots removed, ema.py:19-21) replace `model.*` for the denoiser; here they come from another seed."""
    import lidar_layout_b200 as L
    cfg = C.tiny()
    sd_a, sd_b = random_state_dict(cfg, 0), random_state_dict(cfg, 7)
    sd = dict(sd_a)
    n_shadow = 0
    for k, v in sd_b.items():
        if k.startswith("model.diffusion_model."):
            sd["model_ema." + k[len("model."):].replace(".", "")] = v
            n_shadow += 1
    sd["model_ema.decay"] = torch.tensor(0.9999)
    sd["model_ema.num_updates"] = torch.tensor(12, dtype=torch.int)
    assert n_shadow > 50

    def eps_of(state, use_ema):
        m = L.LatentDiffusion.from_config(tiny_yaml(cfg), use_ema=use_ema)
        m.load_state_dict(state, strict=False)
        g = torch.Generator().manual_seed(5)
        x = torch.randn((2,) + tuple(cfg.latent_shape), generator=g).cuda()
        t = torch.tensor([3, 700], device="cuda")
        with m.ema_scope():
            return m.apply_model(x, t, None)

    e_shadow, e_plain = eps_of(sd, True), eps_of(sd, False)
    assert torch.equal(e_shadow, eps_of(sd_b, False))            # the shadow weights, bit for bit
    assert torch.equal(e_plain, eps_of(sd_a, False))             # use_ema=False ignores them
    assert R.rel_l2(e_shadow.cpu(), e_plain.cpu()) > 0.1         # and the two really differ
