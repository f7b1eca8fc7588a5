"""Developer tool (BASELINE config 4): first-stage autoencoder f_c2_p4 encode / decode throughput sweep on one B200.
    python tests/ae_sweep.py [max_batch]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_layout_b200 import config as C
from lidar_layout_b200.engine import Engine
from lidar_layout_b200.weights import random_encoder_state_dict, random_state_dict

ENC_GF, DEC_GF = 92.6, 119.1     # GFLOP per sample
maxb = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = C.kitti_uncond()
eng = Engine(cfg).load_state_dict({**random_state_dict(cfg, 0), **random_encoder_state_dict(cfg, 0)})


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


B = 1
while B <= maxb:
    x = torch.randn(B, 1, 64, 1024, device="cuda").clamp_(-1, 1)
    z = eng.vq_encode(x)
    reps = max(2, min(20, 256 // B))
    te = timed(lambda: eng.vq_encode(x), reps)
    td = timed(lambda: eng.vq_decode(z), reps)
    print(f"B={B:4d}  encode {te:8.2f} ms ({B / te * 1e3:8.1f} img/s, {B * ENC_GF / te:7.1f} TFLOP/s)   "
          f"decode {td:8.2f} ms ({B / td * 1e3:8.1f} img/s, {B * DEC_GF / td:7.1f} TFLOP/s)   "
          f"round trip {B / (te + td) * 1e3:8.1f} img/s", flush=True)
    del x, z
    B *= 2
