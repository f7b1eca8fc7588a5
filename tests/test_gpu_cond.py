"""GPU (-m gpu): the conditioned U-Nets through the C ABI and the reference-facing Python API against fixtures the
UNMODIFIED reference produced (tests/golden/{tiny_crossattn,tiny_concat,kitti_cam2lidar_L*}.npz):
SpatialTransformer cross-attention, concat / crossattn dispatch (a4), classifier-free guidance (a3).
Tolerances as for the unconditional model (north_star): per-step eps <= 2e-2 relative L2 in bf16, final latents <= 1e-2
... relaxed to 2e-2 under guidance, which amplifies the eps error by the guidance scale (2.5) at every step."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from lidar_layout_b200 import config as C
from lidar_layout_b200.weights import random_state_dict
from oracle import torch_ref as R
from oracle.make_golden import cond_inputs_for, inputs_for

EPS_TOL_BF16 = 2e-2
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return R.rel_l2(a.detach().cpu(), b)


def cfg_for(name):
    if name == "tiny_crossattn":
        return C.tiny(cond="crossattn")
    if name == "tiny_concat":
        return C.tiny(cond="concat")
    return C.kitti_cam2lidar()


@pytest.fixture(scope="module", params=["tiny_crossattn", "tiny_concat", "kitti_cam2lidar_L4", "kitti_cam2lidar_L77"])
def setup(request, built_lib):
    import lidar_layout_b200 as L
    name = request.param
    cfg = cfg_for(name)
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    model = L.LatentDiffusion(cfg, use_ema=False)
    model.load_state_dict(random_state_dict(cfg, 0))
    B, Lc = int(g["B"]), int(g["L"])
    x_T, noise, _ = inputs_for(cfg, B, int(g["S_short"]) + 2)
    cond, uncond = cond_inputs_for(cfg, B, Lc)
    dev = lambda a: None if a is None else torch.from_numpy(a).cuda()
    return name, cfg, g, model, dev(x_T), dev(noise), dev(cond), dev(uncond)


def test_apply_model_with_conditioning(setup):
    name, cfg, g, model, x_T, noise, cond, uncond = setup
    B = x_T.shape[0]
    for tv in (501, 21):
        e = model.apply_model(x_T, torch.full((B,), tv, dtype=torch.long).cuda(), cond)
        assert rel(e, g[f"eps_t{tv}"]) < EPS_TOL_BF16
    # the dict form DiffusionWrapper.forward takes (ddpm.py:2313)
    key = "c_concat" if cfg.conditioning_key == "concat" else "c_crossattn"
    e2 = model.apply_model(x_T, torch.full((B,), 21, dtype=torch.long).cuda(), {key: [cond]})
    assert torch.equal(e, e2)


def test_conditioning_mismatch_is_an_error(setup):
    name, cfg, g, model, x_T, noise, cond, uncond = setup
    t = torch.full((x_T.shape[0],), 5, dtype=torch.long).cuda()
    with pytest.raises(ValueError):
        model.apply_model(x_T, t, None)             # a conditioned model without its conditioning
    bad = cond[:, :, :-1] if cond.dim() == 3 else cond[:, :-1]
    with pytest.raises(ValueError):
        model.apply_model(x_T, t, bad)


def test_conditioned_ddim(setup):
    import lidar_layout_b200 as L
    name, cfg, g, model, x_T, noise, cond, uncond = setup
    if "ddim_eta0_final" not in g.files:
        pytest.skip("eps-only fixture")
    B, S = x_T.shape[0], int(g["S_short"])
    sampler = L.DDIMSampler(model)
    # teacher-forced on the reference's own x_t
    ts = np.flip(R.ddim_schedule(cfg, S, 0.0)[0])
    for i in range(S):
        e = model.apply_model(torch.from_numpy(g["ddim_eta0_xt"][i]).cuda(), torch.full((B,), int(ts[i]), dtype=torch.long).cuda(), cond)
        assert rel(e, g["ddim_eta0_eps"][i]) < EPS_TOL_BF16
    z, inter = sampler.sample(S, batch_size=B, shape=cfg.latent_shape, conditioning=cond, eta=0.0, x_T=x_T.clone())
    assert rel(z, g["ddim_eta0_final"]) < 1e-2
    # the step-by-step path (forced by a callback) agrees with the on-device loop bit for bit
    z2, _ = sampler.sample(S, batch_size=B, shape=cfg.latent_shape, conditioning=cond, eta=0.0, x_T=x_T.clone(),
                           callback=lambda i: None)
    assert torch.equal(z, z2)


def test_classifier_free_guidance(setup):
    import lidar_layout_b200 as L
    name, cfg, g, model, x_T, noise, cond, uncond = setup
    if "ddim_cfg_final" not in g.files:
        pytest.skip("no guidance fixture")
    B, S, scale = x_T.shape[0], int(g["S_short"]), float(g["cfg_scale"])
    sampler = L.DDIMSampler(model)
    t0 = torch.full((B,), int(R.ddim_schedule(cfg, S, 0.0)[0][-1]), dtype=torch.long).cuda()
    e2 = model.apply_model(torch.cat([x_T] * 2), torch.cat([t0] * 2), torch.cat([uncond, cond]))
    assert rel(e2, g["ddim_cfg_eps2b_step0"]) < EPS_TOL_BF16
    # e_u + s (e_c - e_u): exact fp32 arithmetic, op for op
    comb = model.engine.cfg_combine(e2, scale)
    eu, ec = e2.chunk(2)
    assert torch.equal(comb, eu + scale * (ec - eu))
    kw = dict(batch_size=B, shape=cfg.latent_shape, conditioning=cond, unconditional_guidance_scale=scale,
              unconditional_conditioning=uncond)
    z, _ = sampler.sample(S, eta=0.0, x_T=x_T.clone(), **kw)
    assert rel(z, g["ddim_cfg_final"]) < 2e-2
    z2, _ = sampler.sample(S, eta=0.0, x_T=x_T.clone(), callback=lambda i: None, **kw)
    assert torch.equal(z, z2)
    # eta = 1: the sampler draws randn in the reference's order; inject the fixture's noise through the engine call
    ts, tab = R.ddim_schedule(cfg, S, 1.0)
    z1, _ = model.engine.ddim_sample(x_T.clone(), ts, tab, noise=noise[:S], context=cond, uncond_context=uncond,
                                     guidance_scale=scale)
    assert rel(z1, g["ddim_cfg_eta1_final"]) < 2e-2


def test_batch_independence_with_context(setup):
    name, cfg, g, model, x_T, noise, cond, uncond = setup
    if not name.startswith("tiny"):
        pytest.skip("covered on the small configs")
    x3 = torch.cat([x_T, x_T[:1] * 0.5])
    c3 = torch.cat([cond, cond[:1] * -1.0])
    t = torch.tensor([7, 400, 977]).cuda()
    e = model.apply_model(x3, t, c3)
    for i in range(3):
        ei = model.apply_model(x3[i:i + 1], t[i:i + 1], c3[i:i + 1])
        assert torch.equal(ei[0], e[i])
