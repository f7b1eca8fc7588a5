"""Developer tool: DDIM-loop time per U-Net step vs batch size (launch-bound regime at small batches)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_layout_b200 import _lib, config as C, schedule
from lidar_layout_b200.engine import Engine
from lidar_layout_b200.weights import random_state_dict
from oracle import torch_ref as R

cfg = C.kitti_uncond()
eng = Engine(cfg).load_state_dict(random_state_dict(cfg, 0))
ts, tab = R.ddim_schedule(cfg, 50, 0.0)
for B in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64]:
    x = torch.randn(B, 8, 16, 128, device="cuda")
    eng.ddim_sample(x, ts[:5], tab[:5])
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    eng.ddim_sample(x, ts, tab)
    b.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"B={B:3d}: {ms / 50:7.3f} ms per U-Net step (device), host enqueue {1000 * t_host / 50:6.3f} ms per step, "
          f"{(_lib.launch_count() - l0) // 50} launches per step, {B * 50 / ms * 1000 / 50:8.1f} samples/s (DDIM-50 loop only)", flush=True)
