"""CPU, build container only: oracle vs the reference modules executed directly (different seed than the golden
fixtures).  Skipped where /root/reference is absent (e.g. the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")


def test_tiny_model_matches_reference_new_seed():
    from lidar_layout_b200 import config as C
    from lidar_layout_b200.weights import random_state_dict
    from oracle import torch_ref as R
    from oracle.make_golden import build_reference, inputs_for, tiny_yaml
    cfg = C.tiny()
    model = build_reference(cfg, tiny_yaml(cfg))
    sd = random_state_dict(cfg, 7)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected
    x_T, _, z = inputs_for(cfg, 2, 2, seed=77)
    t = torch.tensor([3, 977], dtype=torch.long)      # distinct timesteps per sample
    with torch.no_grad():
        ref = model.apply_model(torch.from_numpy(x_T), t, None)
        got = R.unet_forward(sd, cfg.unet, torch.from_numpy(x_T), t)
        assert R.rel_l2(got, ref) < 1e-5
        refd = model.decode_first_stage(torch.from_numpy(z))
        gotd = R.decode_first_stage(sd, cfg, torch.from_numpy(z))
        assert R.rel_l2(gotd, refd) < 1e-5


def test_state_dict_spec_covers_reference_keys():
    from lidar_layout_b200 import config as C
    from lidar_layout_b200.weights import param_spec
    model, _ = ref_shim.build_reference_lidm()
    cfg = C.kitti_uncond()
    spec = param_spec(cfg)
    ref_sd = model.state_dict()
    for k, (shape, _) in spec.items():
        assert k in ref_sd and tuple(ref_sd[k].shape) == tuple(shape), k
    for k in ref_sd:
        if k.startswith(("model.diffusion_model.", "first_stage_model.decoder.", "first_stage_model.post_quant_conv.",
                         "first_stage_model.quantize.")):
            assert k in spec, k


def test_config_from_reference_yaml():
    import yaml, os
    from lidar_layout_b200 import config as C
    with open(os.path.join(ref_shim.REFERENCE_ROOT, "models/lidm/kitti/uncond/config.yaml")) as f:
        cfg = C.from_reference_dict(yaml.safe_load(f))
    assert cfg == C.kitti_uncond()
