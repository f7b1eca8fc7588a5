"""GPU (-m gpu): the precise ("fp32-class") mode against BASELINE.json north_star's fp32 bars:
per-step eps within 1e-3 relative L2 (teacher-forced on the reference's own x_t), final range images within 1e-2.
Every GEMM of this mode still runs on the tcgen05 kernel (3-way bf16 operand split, fp32 accumulate)."""
import dataclasses

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from lidar_layout_b200 import config as C
from lidar_layout_b200.weights import random_state_dict
from oracle import torch_ref as R
from oracle.make_golden import inputs_for

EPS_TOL_FP32 = 1e-3
IMG_TOL = 1e-2


def rel(a, b):
    return R.rel_l2(a.detach().cpu(), b)


@pytest.fixture(scope="module", params=["tiny", "kitti_uncond"])
def setup(request, built_lib, golden_tiny, golden_kitti):
    from lidar_layout_b200.engine import Engine
    name = request.param
    cfg = dataclasses.replace(C.tiny() if name == "tiny" else C.kitti_uncond(), precision="fp32")
    g = golden_tiny if name == "tiny" else golden_kitti
    sd = random_state_dict(cfg, 0)
    eng = Engine(cfg).load_state_dict(sd)
    return name, cfg, g, eng, sd


def test_eps_within_fp32_bar(setup):
    name, cfg, g, eng, sd = setup
    B = int(g["B"])
    x_T, _, _ = inputs_for(cfg, B, int(g["S_short"]) + 2)
    for tv in (501, 21):
        e = eng.unet_forward(torch.from_numpy(x_T).cuda(), torch.full((B,), tv, dtype=torch.long).cuda())
        assert rel(e, g[f"eps_t{tv}"]) < EPS_TOL_FP32
    for i in range(int(g["S_short"])):
        e = eng.unet_forward(torch.from_numpy(g["ddim_eta0_xt"][i]).cuda(), torch.from_numpy(g["ddim_eta0_t"][i]).cuda())
        assert rel(e, g["ddim_eta0_eps"][i]) < EPS_TOL_FP32


def test_decode_and_final_image_within_bar(setup):
    name, cfg, g, eng, sd = setup
    B, S = int(g["B"]), int(g["S_short"])
    x_T, noise, z = inputs_for(cfg, B, S + 2)
    img, idx = eng.vq_decode(torch.from_numpy(z).cuda(), False, True)
    assert np.array_equal(idx.cpu().numpy(), g["vq_idx"])
    assert rel(img, g["decode_q"]) < IMG_TOL
    assert rel(eng.vq_decode(torch.from_numpy(z).cuda(), True), g["decode_nq"]) < IMG_TOL
    # whole pipeline, free running: x_T -> DDIM (reference schedule) -> decode; checker = oracle decode of the
    # reference's own final latent
    ts, tab = g[f"ddim_S{S}_eta0_timesteps"], g[f"ddim_S{S}_eta0_table"]
    zf, _ = eng.ddim_sample(torch.from_numpy(x_T).cuda(), ts, tab)
    assert rel(zf, g["ddim_eta0_final"]) < IMG_TOL
    ref_img = R.decode_first_stage(sd, cfg, torch.from_numpy(g["ddim_eta0_final"]), force_not_quantize=True)
    assert rel(eng.vq_decode(zf, True), ref_img) < IMG_TOL
    ts, tab = g[f"ddim_S{S}_eta1_timesteps"], g[f"ddim_S{S}_eta1_table"]
    zf1, _ = eng.ddim_sample(torch.from_numpy(x_T).cuda(), ts, tab, noise=torch.from_numpy(noise[:S]).cuda())
    assert rel(zf1, g["ddim_eta1_final"]) < IMG_TOL


def test_ddim50_free_running_tiny(built_lib):
    """DDIM-50 free running on the small config against the oracle run on the host (a few seconds of CPU)."""
    from lidar_layout_b200.engine import Engine
    cfg = dataclasses.replace(C.tiny(), precision="fp32")
    sd = random_state_dict(cfg, 0)
    eng = Engine(cfg).load_state_dict(sd)
    x_T, _, _ = inputs_for(cfg, 1, 1, seed=11)
    ts, tab = R.ddim_schedule(cfg, 50, 0.0)
    zf, _ = eng.ddim_sample(torch.from_numpy(x_T).cuda(), ts, tab)
    z_ref = R.ddim_sample(sd, cfg, 50, torch.from_numpy(x_T), 0.0)
    assert rel(zf, z_ref) < IMG_TOL
    img = eng.vq_decode(zf, True)
    assert rel(img, R.decode_first_stage(sd, cfg, z_ref, force_not_quantize=True)) < IMG_TOL


def test_precise_and_fast_modes_agree_within_bf16_budget(built_lib):
    from lidar_layout_b200.engine import Engine
    cfg = C.tiny()
    sd = random_state_dict(cfg, 0)
    fast = Engine(cfg).load_state_dict(sd)
    prec = Engine(dataclasses.replace(cfg, precision="fp32")).load_state_dict(sd)
    x_T, _, _ = inputs_for(cfg, 2, 1, seed=4)
    t = torch.tensor([100, 900]).cuda()
    a, b = fast.unet_forward(torch.from_numpy(x_T).cuda(), t), prec.unet_forward(torch.from_numpy(x_T).cuda(), t)
    assert rel(a, b.cpu()) < 2e-2


def test_ddim50_free_running_headline_config(built_lib):
    """BASELINE config 2's model (unconditional KITTI-360 LiDM, full size), one sample, all 50 DDIM steps free running,
    then the first-stage decode, against the oracle run on the host (~20 s of CPU): final latent and final range image
    within north_star's 1e-2 in the precise mode; the bf16 path is reported against its own 2e-2-per-step budget."""
    from lidar_layout_b200.engine import Engine
    base = C.kitti_uncond()
    sd = random_state_dict(base, 0)
    x_T, _, _ = inputs_for(base, 1, 1, seed=21)
    ts, tab = R.ddim_schedule(base, 50, 0.0)
    z_ref = R.ddim_sample(sd, base, 50, torch.from_numpy(x_T), 0.0)
    img_ref = R.decode_first_stage(sd, base, z_ref, force_not_quantize=True)
    prec = Engine(dataclasses.replace(base, precision="fp32")).load_state_dict(sd)
    zf, _ = prec.ddim_sample(torch.from_numpy(x_T).cuda(), ts, tab)
    e_lat, e_img = rel(zf, z_ref), rel(prec.vq_decode(zf, True), img_ref)
    del prec
    fast = Engine(base).load_state_dict(sd)
    zb, _ = fast.ddim_sample(torch.from_numpy(x_T).cuda(), ts, tab)
    b_lat, b_img = rel(zb, z_ref), rel(fast.vq_decode(zb, True), img_ref)
    print(f"DDIM-50 full size: precise latent {e_lat:.2e} image {e_img:.2e}; bf16 latent {b_lat:.2e} image {b_img:.2e}")
    # measured on B200: precise 6.8e-5 / 8.2e-5; bf16 U-Net latent 1.2e-3 (a bf16 decoder took the image to 1.06e-2; the
    # default fp16 first stage keeps it within the bar)
    assert e_lat < 1e-3 and e_img < 1e-3
    assert b_lat < IMG_TOL and b_img < IMG_TOL
