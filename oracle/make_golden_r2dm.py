"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/r2dm_{small,full}.npz by running the UNMODIFIED reference EfficientUNet
(lidm/modules/unets/efficient_unet.py) from /root/reference on CPU, loaded STRICTLY (every parameter key and shape) with
the product's seeded state-dict (lidar_layout_b200.weights.random_state_dict); the module's constant buffers (coords,
Fourier frequencies, the 1/sqrt 2 scales) keep the reference's own values.  Runs in the build container only.
    python -m oracle.make_golden_r2dm"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = "/root/reference"


def build_reference(cfg):
    sys.path.insert(0, REF_ROOT)
    from lidm.modules.unets.efficient_unet import EfficientUNet
    from lidar_layout_b200.weights import UNET_PREFIX, random_state_dict
    u = cfg.unet
    net = EfficientUNet(in_channels=u.in_channels, resolution=list(u.image_size), base_channels=u.model_channels,
                        channel_multiplier=list(u.channel_mult), num_residual_blocks=list(u.num_residual_blocks),
                        gn_num_groups=u.gn_num_groups, gn_eps=u.gn_eps, attn_num_heads=u.num_heads,
                        coords_encoding="fourier_features", ring=True).eval()
    sd = {k[len(UNET_PREFIX):]: v for k, v in random_state_dict(cfg, 0).items()}
    params = {k for k, _ in net.named_parameters()}
    assert set(sd) == params, (sorted(set(sd) ^ params))[:10]
    for k, v in net.state_dict().items():
        if k in sd:
            assert tuple(v.shape) == tuple(sd[k].shape), k
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all(k not in params for k in missing), (missing, unexpected)
    return net


def main():
    sys.path.insert(0, ROOT)
    from lidar_layout_b200 import config as C
    torch.set_num_threads(os.cpu_count())
    for name, cfg, B in (("r2dm_small", C.tiny_r2dm(), 2), ("r2dm_full", C.nuscenes_r2dm(), 1)):
        net = build_reference(cfg)
        g = torch.Generator().manual_seed(7)
        x = torch.randn(B, cfg.unet.in_channels, *cfg.unet.image_size, generator=g)
        t = torch.tensor([3, 700][:B])
        with torch.no_grad():
            y = net(x, t)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, x=x.numpy(), t=t.numpy(), eps=y.numpy())
        print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB", "eps rms", float(y.pow(2).mean().sqrt()))


if __name__ == "__main__":
    main()
