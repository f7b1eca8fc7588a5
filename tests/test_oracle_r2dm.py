"""CPU: the oracle of the R2DM pixel-space denoiser (oracle/r2dm_ref.py, SURVEY section 8 f4 / BASELINE config 5) against
outputs of the unmodified reference EfficientUNet (tests/golden/r2dm_*.npz, written by oracle/make_golden_r2dm.py with the
product's seeded weights loaded strictly)."""
import os

import numpy as np
import pytest
import torch

from lidar_layout_b200 import config as C
from lidar_layout_b200.weights import UNET_PREFIX, efficient_unet_param_spec, random_state_dict
from oracle import r2dm_ref as RR

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(cfg, g):
    u = cfg.unet
    sd = {k[len(UNET_PREFIX):]: v for k, v in random_state_dict(cfg, 0).items()}
    return RR.efficient_unet_forward(sd, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), resolution=u.image_size,
                                     base_channels=u.model_channels, channel_multiplier=u.channel_mult,
                                     num_residual_blocks=u.num_residual_blocks, gn_num_groups=u.gn_num_groups,
                                     gn_eps=u.gn_eps, attn_num_heads=u.num_heads)


@pytest.mark.parametrize("name", ["r2dm_small", "r2dm_full"])
def test_efficient_unet_oracle(name):
    cfg = C.tiny_r2dm() if name.endswith("small") else C.nuscenes_r2dm()
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    y = _run(cfg, g)
    err = float((y - torch.from_numpy(g["eps"])).norm() / torch.from_numpy(g["eps"]).norm())
    assert y.shape == g["eps"].shape and err < 1e-5, err


def test_resample_is_the_two_tap_form_the_kernels_use():
    """Resample(up=2) = [x[m-1]/4 + 3x[m]/4, 3x[m]/4 + x[m+1]/4] per axis, Resample(down=2) = [1,3,3,1]/8 at stride 2, ring
    on W and zeros on H (what lidar_layout_b200/csrc/layout.cu's FIR kernels implement)."""
    x = torch.randn(1, 3, 4, 8, generator=torch.Generator().manual_seed(0))
    up = RR.resample(x, up=2)
    xp = torch.nn.functional.pad(torch.nn.functional.pad(x, (1, 1, 0, 0), mode="circular"), (0, 0, 1, 1))
    rows = torch.stack([0.25 * xp[:, :, :-2] + 0.75 * xp[:, :, 1:-1], 0.75 * xp[:, :, 1:-1] + 0.25 * xp[:, :, 2:]], dim=3)
    rows = rows.reshape(1, 3, 8, 10)
    want = torch.stack([0.25 * rows[..., :-2] + 0.75 * rows[..., 1:-1], 0.75 * rows[..., 1:-1] + 0.25 * rows[..., 2:]], dim=4)
    assert torch.allclose(up, want.reshape(1, 3, 8, 16), atol=1e-6)
    down = RR.resample(x, down=2)
    k = torch.tensor([1., 3., 3., 1.]) / 8
    k2 = (k[:, None] * k[None, :])[None, None].repeat(3, 1, 1, 1)
    assert torch.allclose(down, torch.nn.functional.conv2d(xp, k2, stride=2, groups=3), atol=1e-6)


def test_param_spec_counts():
    spec = efficient_unet_param_spec(C.nuscenes_r2dm().unet)
    n = sum(int(np.prod(s)) for s, _ in spec.values())
    assert 30e6 < n < 40e6          # the R2DM U-Net: 31 M parameters
    cfg = C.from_reference_dict({"model": {"target": "lidm.models.diffusion.ddpm_r2dm.R2DMDiffusion", "params": {
        "timesteps": 1024, "linear_start": 0.0015, "linear_end": 0.0195, "image_size": [32, 1024], "channels": 2,
        "unet_config": {"target": "lidm.modules.unets.efficient_unet.EfficientUNet", "params": {
            "in_channels": 2, "resolution": [32, 1024], "base_channels": 64, "temb_channels": None,
            "channel_multiplier": [1, 2, 4, 8], "num_residual_blocks": [3, 3, 3, 3], "gn_num_groups": 8, "gn_eps": 1e-6,
            "attn_num_heads": 8, "coords_encoding": "fourier_features", "ring": True}},
        "cond_stage_config": "__is_unconditional__"}}})
    assert cfg.unet == C.nuscenes_r2dm().unet and cfg.timesteps == 1024
