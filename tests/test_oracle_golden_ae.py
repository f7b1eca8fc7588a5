"""CPU: the oracle's first-stage encoder (Encoder.forward, VQModelInterface.encode) against
fixtures produced by the UNMODIFIED reference (`python -m oracle.make_golden --ae`)."""
import os

import numpy as np
import pytest
import torch

from lidar_layout_b200 import config as C
from lidar_layout_b200.weights import random_encoder_state_dict, random_state_dict
from oracle import torch_ref as R
from oracle.make_golden import ae_images_for, sd_digest

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5


@pytest.mark.parametrize("name", ["tiny_ae", "kitti_ae"])
def test_encode_and_round_trip(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = C.tiny() if name == "tiny_ae" else C.kitti_uncond()
    sd = {**random_state_dict(cfg, 0), **random_encoder_state_dict(cfg, 0)}
    assert sd_digest(sd) == bytes(g["weights_digest"]).decode()
    x = torch.from_numpy(ae_images_for(cfg, int(g["B"])))
    z = R.encode_first_stage(sd, cfg, x)
    assert R.rel_l2(z, g["encode"]) < TOL
    assert R.rel_l2(R.get_first_stage_encoding(cfg, z), g["encoding_scaled"]) < TOL
    assert R.rel_l2(R.decode_first_stage(sd, cfg, torch.from_numpy(g["encode"]), force_not_quantize=True), g["recon_nq"]) < TOL
    assert R.rel_l2(R.decode_first_stage(sd, cfg, torch.from_numpy(g["encode"])), g["recon_q"]) < TOL
