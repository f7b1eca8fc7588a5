"""CPU: host-side logic of the product package (no GPU compute): schedules against the reference fixtures, config
parsing, weight spec, the C-ABI library (loads, exports every symbol include/*.h declares, fails loudly without a
GPU), and the DDIM sampler's loop-segmentation logic against a fake engine."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from lidar_layout_b200 import _lib, config as C, schedule
from lidar_layout_b200.weights import param_spec, random_state_dict, unet_blocks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["tiny", "kitti_uncond"])
def test_product_schedule_matches_reference_fixture(name, golden_tiny, golden_kitti):
    g = golden_tiny if name == "tiny" else golden_kitti
    cfg = C.tiny() if name == "tiny" else C.kitti_uncond()
    bufs = schedule.ddpm_buffers(cfg.beta_schedule, cfg.timesteps, cfg.linear_start, cfg.linear_end)
    S = int(g["S_short"])
    for S_, eta in ((50, 0.0), (50, 1.0), (S, 0.0), (S, 1.0)):
        ts = schedule.make_ddim_timesteps("uniform", S_, cfg.timesteps)
        tab, _ = schedule.ddim_table(bufs["alphas_cumprod"], ts, eta)
        assert np.array_equal(ts, g[f"ddim_S{S_}_eta{int(eta)}_timesteps"])
        assert np.array_equal(tab, g[f"ddim_S{S_}_eta{int(eta)}_table"])


def test_ddim_timestep_counts():
    # "DDIM-50" is 50 steps only when T=1000; T=1024 gives 52
    assert len(schedule.make_ddim_timesteps("uniform", 50, 1000)) == 50
    assert len(schedule.make_ddim_timesteps("uniform", 50, 1024)) == 52
    assert len(schedule.make_ddim_timesteps("uniform", 256, 1024)) == 256


def test_config_roundtrip_from_reference_format():
    from oracle.make_golden import tiny_yaml
    cfg = C.tiny()
    assert C.from_reference_dict(tiny_yaml(cfg)).unet == cfg.unet
    assert C.from_reference_dict(tiny_yaml(cfg)).ae == cfg.ae
    bad = tiny_yaml(cfg)
    bad["model"]["target"] = "lidm.models.diffusion.ddpm_r2dm.R2DMDiffusion"
    with pytest.raises(ValueError):
        C.from_reference_dict(bad)


def test_unet_topology_of_released_config():
    inputs, middle, outputs, ch = unet_blocks(C.kitti_uncond().unet)
    assert len(inputs) == 9 and len(outputs) == 9 and ch == 256
    n_res = sum(1 for blk in inputs + [middle] + outputs for l in blk if l[0] == "res")
    n_attn = sum(1 for blk in inputs + [middle] + outputs for l in blk if l[0] == "attn")
    assert (n_res, n_attn) == (17, 16)
    spec = param_spec(C.kitti_uncond())
    n_unet = sum(int(np.prod(s)) for k, (s, _) in spec.items() if k.startswith("model.diffusion_model."))
    assert n_unet == 257_748_232 or abs(n_unet - 257.75e6) < 0.01e6


def test_random_state_dict_is_deterministic_and_nondegenerate():
    cfg = C.tiny()
    a, b = random_state_dict(cfg, 3, as_torch=False), random_state_dict(cfg, 3, as_torch=False)
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert all(np.abs(v).max() > 0 for v in a.values())


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lidm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lidm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/lidm_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built_lib):
    lib = _lib.load()
    cc = _lib.CConfig()
    h = ctypes.c_void_p()
    rc = lib.lidm_create(ctypes.byref(cc), ctypes.byref(h))
    assert rc == -2 and b"no CPU fallback" in lib.lidm_last_error(None)
    import lidar_layout_b200 as L
    with pytest.raises(L.LidmError):
        L.LatentDiffusion(C.tiny())
    with pytest.raises((ValueError, L.LidmError)):
        L.ops.ddim_step(torch.zeros(4), torch.zeros(4), (0.5, 0.6, 0.0, 0.7))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lidar_layout_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/", "ORACLEDOC"), f"{f} references the oracle"


class _FakeEngine:
    """Records lidm_ddim_sample segments; x advances by +1 per step so order/coverage are checkable."""

    def __init__(self):
        self.calls = []

    def ddim_sample(self, x, timesteps, table, noise=None, temperature=1.0, want_pred_x0=False):
        self.calls.append((list(timesteps), None if noise is None else noise.shape[0]))
        return x + len(timesteps), x * 0 + len(self.calls)


class _FakeModel:
    def __init__(self, cfg):
        bufs = schedule.ddpm_buffers(cfg.beta_schedule, cfg.timesteps, cfg.linear_start, cfg.linear_end)
        self.num_timesteps = cfg.timesteps
        self.device = torch.device("cpu")
        self.engine = _FakeEngine()
        for k, v in bufs.items():
            setattr(self, k, v)


@pytest.mark.parametrize("S,log_every_t,eta", [(50, 100, 0.0), (50, 10, 1.0), (4, 1, 0.0)])
def test_ddim_sampler_segments_cover_loop_in_reference_order(S, log_every_t, eta):
    from lidar_layout_b200.ddim import DDIMSampler
    cfg = C.kitti_uncond()
    m = _FakeModel(cfg)
    s = DDIMSampler(m)
    x_T = torch.zeros(2, 8, 16, 128)
    out, inter = s.sample(S, 2, (8, 16, 128), eta=eta, x_T=x_T, log_every_t=log_every_t)
    n = len(s.ddim_timesteps)
    assert float(out[0, 0, 0, 0]) == n                                   # every step ran exactly once
    flat = [t for seg, _ in m.engine.calls for t in reversed(seg)]       # engine walks each segment backwards
    assert flat == list(np.flip(s.ddim_timesteps))
    expected_logs = sum(1 for idx in range(n) if idx % log_every_t == 0 or idx == n - 1)
    assert len(inter["x_inter"]) == 1 + expected_logs == len(inter["pred_x0"])
    for seg, nz in m.engine.calls:
        assert (nz == len(seg)) if eta > 0 else (nz is None)
