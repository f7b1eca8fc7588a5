"""Developer tool: U-Net step time and first-stage decode / encode time per numeric mode at batch B (one B200).
    python tests/mode_sweep.py [B]"""
import dataclasses, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_layout_b200 import config as C
from lidar_layout_b200.engine import Engine
from lidar_layout_b200.weights import random_encoder_state_dict, random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64


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


base = C.kitti_uncond()
sd = {**random_state_dict(base, 0), **random_encoder_state_dict(base, 0)}
x = torch.randn(B, 8, 16, 128, device="cuda")
t = torch.full((B,), 501, dtype=torch.long, device="cuda")
img = torch.randn(B, 1, 64, 1024, device="cuda").clamp_(-1, 1)
for unet, ae in (("bf16", "bf16"), ("bf16", "fp16"), ("fp16", "fp16"), ("bf16", "fp32"), ("fp32", "fp32")):
    eng = Engine(dataclasses.replace(base, precision=unet, ae_precision=ae)).load_state_dict(sd)
    tu = timed(lambda: eng.unet_forward(x, t), 5)
    td = timed(lambda: eng.vq_decode(x), 3)
    te = timed(lambda: eng.vq_encode(img), 3)
    print(f"B={B} unet={unet:5s} ae={ae:5s}: unet {tu:8.2f} ms/step   decode {td:8.2f} ms   encode {te:8.2f} ms", flush=True)
    del eng
    torch.cuda.empty_cache()
