"""Developer probe: run each GPU parity check in its own subprocess (a CUDA fault poisons the context) and print
relative errors instead of asserting.  Usage on a GPU box:  python tests/gpu_probe.py [check ...]"""
from __future__ import annotations

import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    import torch
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_backproject():
    import numpy as np, torch
    from lidar_layout_b200 import ops
    from oracle import torch_ref as R
    g = np.load(os.path.join(ROOT, "tests/golden/kitti_uncond.npz"))
    img = torch.from_numpy(g["bp_img"]).cuda()[None]
    xyz, mask = ops.backproject(img, (3, -25), (1.0, 56.0), 5.84, True)
    ref = torch.from_numpy(g["bp_xyz"])
    print("backproject max abs err", float((xyz[0].cpu().double() - ref).abs().max()), "mask sum", int(mask.sum()),
          "ref pts", g["bp_pcd"].shape[0])


def check_ddim():
    import torch
    from lidar_layout_b200 import ops
    from oracle import torch_ref as R
    torch.manual_seed(0)
    x, e, n = torch.randn(2, 8, 16, 128), torch.randn(2, 8, 16, 128), torch.randn(2, 8, 16, 128)
    coef = (0.5123, 0.6234, 0.1, (1 - 0.5123) ** 0.5)
    xp, x0 = ops.ddim_step(x.cuda(), e.cuda(), coef, n.cuda(), 0.9)
    import numpy as np
    rp, r0 = R.ddim_step(x, e, np.float32(coef), n, 0.9)
    print("ddim x_prev bit-exact", bool((xp.cpu() == rp).all()), "pred_x0 bit-exact", bool((x0.cpu() == r0).all()),
          rel(xp, rp))


def check_gn():
    import torch, torch.nn.functional as F
    from lidar_layout_b200 import ops
    torch.manual_seed(0)
    for (B, C, H, W) in [(2, 256, 16, 128), (1, 768, 16, 128), (2, 64, 8, 64), (2, 128, 4, 32), (1, 1536, 8, 64)]:
        x = torch.randn(B, C, H, W) * 2 + 0.3
        ga, be = torch.randn(C), torch.randn(C)
        for silu in (False, True):
            y = ops.group_norm(x.cuda(), ga.cuda(), be.cuda(), 1e-5, 32, silu)
            xr = x.bfloat16().float()
            r = F.group_norm(xr, 32, ga, be, 1e-5)
            if silu:
                r = F.silu(r)
            print(f"gn {B,C,H,W} silu={silu} rel {rel(y, r):.2e}")


def check_conv():
    import torch
    from lidar_layout_b200 import ops
    from oracle import torch_ref as R
    torch.manual_seed(0)
    cases = [
        # B, Cin, H, W, Cout, kh, kw, pad, stride
        (1, 64, 4, 32, 128, 1, 1, (0, 0, 0, 0), 1),
        (2, 128, 16, 128, 128, 1, 1, (0, 0, 0, 0), 1),
        (2, 256, 16, 128, 256, 3, 3, (1, 1, 1, 1), 1),
        (2, 512, 8, 64, 512, 3, 3, (1, 1, 1, 1), 1),
        (2, 1024, 4, 32, 1024, 3, 3, (1, 1, 1, 1), 1),
        (2, 8, 16, 128, 256, 3, 3, (1, 1, 1, 1), 1),
        (2, 256, 16, 128, 8, 3, 3, (1, 1, 1, 1), 1),
        (2, 256, 16, 128, 256, 3, 3, (1, 1, 1, 1), 2),
        (1, 128, 8, 256, 128, 1, 4, (1, 2, 0, 0), 1),
        (1, 128, 8, 256, 128, 1, 5, (2, 2, 0, 0), 1),
        (1, 64, 8, 256, 1, 1, 4, (1, 2, 0, 0), 1),
        (1, 128, 8, 256, 64, 1, 4, (1, 2, 0, 0), 1),
    ]
    for (B, Cin, H, W, Cout, kh, kw, pad, stride) in cases:
        x = torch.randn(B, Cin, H, W)
        w = torch.randn(Cout, Cin, kh, kw) / (Cin * kh * kw) ** 0.5
        b = torch.randn(Cout)
        t0 = time.time()
        y = ops.circular_conv2d(x.cuda(), w.cuda(), b.cuda(), pad, stride)
        ref = R.circular_conv2d(x.bfloat16().float(), w.bfloat16().float(), b, pad, stride)
        print(f"conv B{B} {Cin}->{Cout} {H}x{W} k{kh}x{kw} s{stride} rel {rel(y, ref):.2e}  ({time.time()-t0:.2f}s)")
    # residual
    x = torch.randn(2, 256, 16, 128); w = torch.randn(256, 256, 3, 3) / 48; b = torch.randn(256)
    res = torch.randn(2, 256, 16, 128)
    y = ops.circular_conv2d(x.cuda(), w.cuda(), b.cuda(), (1, 1, 1, 1), 1, residual=res.cuda())
    ref = R.circular_conv2d(x.bfloat16().float(), w.bfloat16().float(), b, (1, 1, 1, 1)) + res.bfloat16().float()
    print(f"conv+residual rel {rel(y, ref):.2e}")


def check_attn():
    import torch
    from lidar_layout_b200 import ops
    from oracle import torch_ref as R
    torch.manual_seed(0)
    for (B, heads, T) in [(1, 2, 128), (2, 4, 512), (1, 8, 2048), (2, 32, 128)]:
        qkv = torch.randn(B, heads * 96, T)
        y = ops.qkv_attention_legacy(qkv.cuda(), heads)
        ref = R.qkv_attention_legacy(qkv.bfloat16().float(), heads)
        print(f"attn B{B} heads{heads} T{T} rel {rel(y, ref):.2e}")


def _model(name):
    import numpy as np, torch, dataclasses
    from lidar_layout_b200 import config as C
    from lidar_layout_b200.engine import Engine
    from lidar_layout_b200.weights import random_state_dict
    precise = name.endswith("_p")
    name = name[:-2] if precise else name
    cfg = C.tiny() if name == "tiny" else C.kitti_uncond()
    if precise:
        cfg = dataclasses.replace(cfg, precision="fp32")
    g = np.load(os.path.join(ROOT, f"tests/golden/{name}.npz"))
    name = name + ("_p" if precise else "")
    t0 = time.time()
    sd = random_state_dict(cfg, 0)
    eng = Engine(cfg).load_state_dict(sd)
    torch.cuda.synchronize()
    print(f"[{name}] engine ready in {time.time()-t0:.1f}s")
    return cfg, g, eng, sd


def _check_model(name):
    import numpy as np, torch
    from oracle.make_golden import inputs_for
    cfg, g, eng, sd = _model(name)
    name = name + "" 
    B, S = int(g["B"]), int(g["S_short"])
    x_T, noise, z = inputs_for(cfg, B, S + 2)
    for tv in (501, 21):
        e = eng.unet_forward(torch.from_numpy(x_T).cuda(), torch.full((B,), tv, dtype=torch.long).cuda())
        print(f"[{name}] eps t={tv} rel {rel(e, torch.from_numpy(g[f'eps_t{tv}'])):.3e}")
    # teacher-forced per-step eps
    for i in range(S):
        e = eng.unet_forward(torch.from_numpy(g["ddim_eta0_xt"][i]).cuda(), torch.from_numpy(g["ddim_eta0_t"][i]).cuda())
        print(f"[{name}] teacher-forced step {i} t={int(g['ddim_eta0_t'][i][0])} eps rel {rel(e, torch.from_numpy(g['ddim_eta0_eps'][i])):.3e}")
    ts, tab = g[f"ddim_S{S}_eta0_timesteps"], g[f"ddim_S{S}_eta0_table"]
    xf, _ = eng.ddim_sample(torch.from_numpy(x_T).cuda(), ts, tab)
    print(f"[{name}] ddim eta0 final rel {rel(xf, torch.from_numpy(g['ddim_eta0_final'])):.3e}")
    ts, tab = g[f"ddim_S{S}_eta1_timesteps"], g[f"ddim_S{S}_eta1_table"]
    xf, _ = eng.ddim_sample(torch.from_numpy(x_T).cuda(), ts, tab, noise=torch.from_numpy(noise[:S]).cuda())
    print(f"[{name}] ddim eta1 final rel {rel(xf, torch.from_numpy(g['ddim_eta1_final'])):.3e}")
    img, idx = eng.vq_decode(torch.from_numpy(z).cuda(), False, True)
    print(f"[{name}] decode q rel {rel(img, torch.from_numpy(g['decode_q'])):.3e} idx match {(idx.cpu().numpy() == g['vq_idx']).mean():.5f}")
    img = eng.vq_decode(torch.from_numpy(z).cuda(), True)
    print(f"[{name}] decode nq rel {rel(img, torch.from_numpy(g['decode_nq'])):.3e}")


def check_tiny():
    _check_model("tiny")


def check_kitti():
    _check_model("kitti_uncond")


def check_tiny_p():
    _check_model("tiny_p")


def check_kitti_p():
    _check_model("kitti_uncond_p")


CHECKS = ["backproject", "ddim", "gn", "conv", "attn", "tiny", "kitti"]

if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        globals()["check_" + sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or CHECKS
    for n in names:
        print(f"===== {n}", flush=True)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", n], timeout=600, capture_output=True,
                               text=True)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print(f"!!! {n} exit {r.returncode}\n{r.stderr[-3000:]}")
        except subprocess.TimeoutExpired as e:
            print(f"!!! {n} TIMEOUT", (e.stdout or b"")[-2000:], (e.stderr or b"")[-2000:])
        print(f"===== {n} done in {time.time()-t0:.1f}s", flush=True)
