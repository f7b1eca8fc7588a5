"""GPU (-m gpu): the numeric modes of the tensor-core paths (lidm_config.precision / ae_precision) against the
reference fixtures.  north_star: per-step eps within 2e-2 (bf16) / 1e-3 (fp32), final range image within 1e-2.
  U-Net:       bf16 (benchmarked), fp32 = 3-way bf16 operand split, fp16 = IEEE half at the bf16 tensor rate
  first stage: fp16 (benchmarked default under a bf16 U-Net), fp32 = operand split, bf16 (legacy: misses the 1e-2 bar)"""
import dataclasses
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from lidar_layout_b200 import config as C
from lidar_layout_b200.weights import random_encoder_state_dict, random_state_dict
from oracle import torch_ref as R
from oracle.make_golden import ae_images_for, inputs_for

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
IMG_TOL = 1e-2


def rel(a, b):
    return R.rel_l2(a.detach().cpu(), b)


def _cfg(name, **kw):
    return dataclasses.replace(C.tiny() if name.startswith("tiny") else C.kitti_uncond(), **kw)


@pytest.mark.parametrize("name", ["tiny", "kitti_uncond"])
@pytest.mark.parametrize("ae_mode,tol", [("fp16", IMG_TOL), ("fp32", 2e-3), ("bf16", 4e-2)])
def test_first_stage_modes(built_lib, name, ae_mode, tol):
    """Decoder and encoder of every first-stage mode against the reference; `bf16` is kept as the documented legacy
    mode (it does NOT meet the 1e-2 bar: that is why it is no longer the default)."""
    import lidar_layout_b200 as L
    cfg = _cfg(name, ae_precision=ae_mode)
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    ga = np.load(os.path.join(GOLDEN_DIR, name.replace("_uncond", "") + "_ae.npz"))
    model = L.LatentDiffusion(cfg, use_ema=False)
    model.load_state_dict({**random_state_dict(cfg, 0), **random_encoder_state_dict(cfg, 0)})
    _, _, z = inputs_for(cfg, int(g["B"]), int(g["S_short"]) + 2)
    img, idx = model.engine.vq_decode(torch.from_numpy(z).cuda(), False, True)
    assert np.array_equal(idx.cpu().numpy(), g["vq_idx"])
    e_q = rel(img, g["decode_q"])
    e_nq = rel(model.engine.vq_decode(torch.from_numpy(z).cuda(), True), g["decode_nq"])
    x = torch.from_numpy(ae_images_for(cfg, int(ga["B"]))).cuda()
    zz = model.encode_first_stage(x)
    e_enc = rel(zz, ga["encode"])
    e_rt = rel(model.decode_first_stage(zz, force_not_quantize=True), ga["recon_nq"])
    print(f"[{name} ae={ae_mode}] decode {e_q:.3e} / {e_nq:.3e}  encode {e_enc:.3e}  round trip {e_rt:.3e}")
    assert e_q < tol and e_nq < tol and e_enc < tol and e_rt < 2 * tol
    assert torch.equal(model.encode_first_stage(x[:1])[0], zz[0])          # batch-invariant in every mode


@pytest.mark.parametrize("name", ["tiny", "kitti_uncond"])
def test_unet_fp16_mode(built_lib, name):
    """IEEE-half U-Net (same kernels, same tensor rate as bf16): teacher-forced eps against the reference fixture.
    Three more mantissa bits than bf16 (measured 6.8e-3 there)."""
    from lidar_layout_b200.engine import Engine
    cfg = _cfg(name, precision="fp16")
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    eng = Engine(cfg).load_state_dict(random_state_dict(cfg, 0))
    errs = []
    for i in range(int(g["S_short"])):
        e = eng.unet_forward(torch.from_numpy(g["ddim_eta0_xt"][i]).cuda(), torch.from_numpy(g["ddim_eta0_t"][i]).cuda())
        errs.append(rel(e, g["ddim_eta0_eps"][i]))
    print(f"[{name} unet=fp16] teacher-forced eps rel {['%.2e' % v for v in errs]}")
    assert max(errs) < 2.5e-3
    B, S = int(g["B"]), int(g["S_short"])
    x_T, _, _ = inputs_for(cfg, B, S + 2)
    ts, tab = g[f"ddim_S{S}_eta0_timesteps"], g[f"ddim_S{S}_eta0_table"]
    xf, _ = eng.ddim_sample(torch.from_numpy(x_T).cuda(), ts, tab)
    assert rel(xf, g["ddim_eta0_final"]) < 2e-3


def test_mode_validation(built_lib):
    from lidar_layout_b200.engine import Engine
    with pytest.raises(ValueError):
        Engine(dataclasses.replace(C.tiny(), precision="fp64"))
    with pytest.raises(ValueError):
        Engine(dataclasses.replace(C.tiny(), ae_precision="int8"))
    assert C.tiny().ae_precision_resolved == "fp16"
    assert dataclasses.replace(C.tiny(), precision="fp32").ae_precision_resolved == "fp32"
