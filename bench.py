#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): samples/sec for DDIM-50 LiDM sampling of 64x1024 range images
(unconditional KITTI-360 config, random-init weights, synthetic noise) -> VQ decode -> back-projection.

    python bench.py --gpus N --steps K --warmup W             # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU (oracle port)

One "step" = one full pass of the hot path over one batch of B samples per GPU: 50 U-Net evaluations with the
fused DDIM update, first-stage decode, back-projection of all B images (and, for N>1, the single all-gather of the
range images).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

UNET_GFLOP_PER_SAMPLE_STEP = 171.706      # BASELINE.md section 2 (FlopCounterMode on the reference module)
SAMPLE_GFLOP = 8704.9                     # 50 x U-Net + VQ + decoder
DDIM_STEPS = 50
METRIC = "samples/sec DDIM-50 64x1024 LiDM (sampling + VQ decode + back-projection)"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def captured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
    `ncu --set full` capture (profiles/roofline_traffic.json); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(p) as f:
            return float(json.load(f)["traffic_bytes_per_launch_mean"])
    except Exception:
        return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


WORKLOADS = {
    # name: (config factory name, description, GFLOP per sample: 50 U-Net evaluations + VQ + decoder)
    "uncond": ("kitti_uncond", "unconditional LiDM KITTI-360 64x1024 (models/lidm/kitti/uncond, f_c2_p4 AE, random-init), "
                               "DDIM-50 eta=0 + VQ decode + back-projection", SAMPLE_GFLOP),
    # BASELINE config 3(A): the cross-attention conditioned U-Net (SpatialTransformer blocks), synthetic (B,4,512) context
    "cam2lidar": ("kitti_cam2lidar", "cross-attention conditioned LiDM KITTI-360 64x1024 (models/lidm/kitti/cam2lidar, "
                                     "SpatialTransformer U-Net, context (B,4,512) synthetic, random-init), DDIM-50 eta=0 + "
                                     "VQ decode + back-projection", 50 * 240.6 + 0.537 + 119.06),
}


def workload_config(B, world, name="uncond"):
    return {"workload": WORKLOADS[name][1],
            "batch_per_gpu": B, "global_batch": B * world, "ddim_steps": DDIM_STEPS,
            "l2": "working set per step (GBs of activations + 0.5 GB weights) >> 126 MB L2",
            "parallelism": f"batch-sharded x{world}, one all-gather of range images"}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(cfg, sd, n_unet_steps, B, threads):
    """The reference algorithm (oracle port, plain torch fp32 on the host): DDIM steps + decode + back-projection.
    Returns seconds for (n_unet_steps U-Net evaluations incl. DDIM update, decode, back-projection) at batch B."""
    from oracle import torch_ref as R
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1000)
    x = torch.randn((B,) + tuple(cfg.latent_shape), generator=g)
    ts, table = R.ddim_schedule(cfg, DDIM_STEPS, 0.0)
    t0 = time.perf_counter()
    n = len(ts)
    for i, step in enumerate(np.flip(ts)[:n_unet_steps]):
        t = torch.full((B,), int(step), dtype=torch.long)
        e = R.unet_forward(sd, cfg.unet, x, t)
        x, _ = R.ddim_step(x, e, table[n - 1 - i])
    t_unet = time.perf_counter() - t0
    t0 = time.perf_counter()
    img = R.decode_first_stage(sd, cfg, x)
    t_dec = time.perf_counter() - t0
    ds = dict(fov=cfg.dataset.fov, depth_range=cfg.dataset.depth_range, depth_scale=cfg.dataset.depth_scale,
              log_scale=cfg.dataset.log_scale)
    t0 = time.perf_counter()
    for b in range(B):
        R.range2pcd(R.custom_to_unit(img[b, 0].numpy()), **ds)
    t_bp = time.perf_counter() - t0
    return t_unet, t_dec, t_bp


def cpu_samples_per_sec(cfg, sd, n_unet_steps, B, threads):
    t_unet, t_dec, t_bp = cpu_reference_run(cfg, sd, n_unet_steps, B, threads)
    full = t_unet * (DDIM_STEPS / n_unet_steps) + t_dec + t_bp        # extrapolate the loop to 50 steps
    return B / full, dict(t_unet_per_step=t_unet / n_unet_steps, t_decode=t_dec, t_backproject=t_bp)


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from lidar_layout_b200 import config as C
    from lidar_layout_b200.weights import random_state_dict
    cfg = C.kitti_uncond()
    sd = random_state_dict(cfg, 0)
    threads = os.cpu_count() or 1
    B, n_steps = 2, 10        # bounded sample: ~6 s of host work per repeat on 16 cores
    vals = []
    for i in range(args.warmup + args.steps):
        if i >= min(args.warmup, 1) + args.steps and vals:      # bounded: the CPU arm is slow
            break
        v, detail = cpu_samples_per_sec(cfg, sd, n_steps, B, threads)
        if i >= min(args.warmup, 1):
            vals.append(v)
        if len(vals) >= min(args.steps, 3):
            break
    v = float(np.mean(vals))
    sample = (f"B={B}, {n_steps} of {DDIM_STEPS} DDIM steps timed and extrapolated x{DDIM_STEPS // n_steps}, "
              f"full decode + back-projection; oracle port (torch fp32 CPU), {len(vals)} repeats")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": min(args.warmup, 1), "ms_per_step": 1000.0 * B / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args.batch, 1),
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(args):
    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import lidar_layout_b200 as L
    from lidar_layout_b200 import _lib, config as C, parallel
    from lidar_layout_b200.weights import random_state_dict

    cfg = getattr(C, WORKLOADS[args.workload][0])()
    B = args.batch
    sd = random_state_dict(cfg, 0)
    model = L.LatentDiffusion(cfg, device=dev, use_ema=False, precision=args.precision)
    model.load_state_dict(sd)
    sampler = L.DDIMSampler(model)
    sampler.make_schedule(DDIM_STEPS, ddim_eta=0.0)
    ts, table = sampler.ddim_timesteps, sampler.ddim_table
    assert len(ts) == DDIM_STEPS
    ds = cfg.dataset
    global_B = B * world
    x_T_global, _ = parallel.global_noise((global_B,) + tuple(cfg.latent_shape), seed=1000)
    x_T_host = parallel.local_slice(x_T_global, rank, world).pin_memory()
    x_T_dev = x_T_host.to(dev)
    eng = model.engine
    cond_kw, cond = {}, None
    if cfg.conditioning_key == "crossattn":
        gctx = torch.Generator().manual_seed(1007)
        ctx_global = torch.randn((global_B, 4, cfg.unet.context_dim), generator=gctx)
        cond = parallel.local_slice(ctx_global, rank, world).to(dev)
        cond_kw = dict(context=cond)

    def device_step():
        z, _ = eng.ddim_sample(x_T_dev, ts, table, **cond_kw)
        img = eng.vq_decode(z)
        xyz, mask = L.ops.backproject(img, ds.fov, ds.depth_range, ds.depth_scale, ds.log_scale)
        if world > 1:
            img = parallel.all_gather_batch(img, global_B)
        return img, xyz, mask

    xyz_host = torch.empty((B, 3, ds.size[0], ds.size[1]), dtype=torch.float32).pin_memory()
    mask_host = torch.empty((B, ds.size[0], ds.size[1]), dtype=torch.uint8).pin_memory()

    def e2e_step():
        # the call sequence a reference user makes (scripts/sample.py:89-110,129), host buffers in and out
        x = x_T_host.to(dev, non_blocking=True)
        with model.ema_scope("Plotting"):
            z, _ = sampler.sample(DDIM_STEPS, batch_size=B, shape=cfg.latent_shape, eta=0.0, x_T=x, conditioning=cond)
        img = model.decode_first_stage(z)
        xyz, mask = L.ops.backproject(img, ds.fov, ds.depth_range, ds.depth_scale, ds.log_scale)
        if world > 1:
            img = parallel.all_gather_batch(img, global_B)
        xyz_host.copy_(xyz, non_blocking=True)
        mask_host.copy_(mask, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        device_step()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    ms = timed(device_step, args.steps)
    launches = _lib.launch_count() - l0
    clk = clocks.stop() if rank == 0 else {}
    value = global_B * args.steps / (ms / 1000.0)

    # end to end through the reference-facing API, host <-> device copies inside the timed region
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = global_B * args.steps / (ms_e2e / 1000.0)

    # per-kernel-class device time for the roofline line: one extra step bracketed by CUDA events per op
    roof = None
    unet_ms = None
    if rank == 0:
        torch.cuda.synchronize()
        _lib.profile_begin()
        z, _ = eng.ddim_sample(x_T_dev, ts, table, **cond_kw)
        prof_unet = _lib.profile_end()
        _lib.profile_begin()
        img = eng.vq_decode(z)
        prof_dec = _lib.profile_end()
        peaks = measured_peaks()
        g = prof_unet["conv_gemm"]
        ach = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        unet_ms = sum(v["ms"] for v in prof_unet.values()) / DDIM_STEPS
        roof = {
            "bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM conv, all U-Net launches of one DDIM-50 loop)",
            "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": ach / peaks["bf16_tflops_sustained"], "peak_source": peaks["source"] + " (sustained: kernel timed inside a long step)",
            "traffic": captured_traffic(),
            "traffic_unit": "bytes per launch (mean of the 12 launches in profiles/roofline_traffic.json)",
            "launches": g["launches"], "avg_launch_ms": g["ms"] / max(g["launches"], 1),
            "share_of_step": {k: v["ms"] / max(sum(x["ms"] for x in prof_unet.values()), 1e-9) for k, v in prof_unet.items()},
            "groupnorm_GBps": (prof_unet["groupnorm"]["bytes"] / (prof_unet["groupnorm"]["ms"] * 1e-3) / 1e9
                               if prof_unet["groupnorm"]["ms"] > 0 else None),
            "attention_TFLOPs": (prof_unet["attention"]["flops"] / (prof_unet["attention"]["ms"] * 1e-3) / 1e12
                                 if prof_unet["attention"]["ms"] > 0 else None),
            "decoder_ms": sum(v["ms"] for v in prof_dec.values()),
            "whole_pipeline_frac_of_tensor_peak": value / world * WORKLOADS[args.workload][2] / 1e3 / peaks["bf16_tflops_sustained"],
        }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "uncond":
        threads = os.cpu_count() or 1
        v, detail = cpu_samples_per_sec(cfg, sd, 10, 2, threads)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"B=2, 10 of {DDIM_STEPS} DDIM steps timed and extrapolated x5, full decode + back-projection "
                         f"(oracle port, torch fp32); per-step {detail['t_unet_per_step']:.2f}s decode {detail['t_decode']:.2f}s"}

    if rank == 0:
        h2d = x_T_host.numel() * 4
        d2h = xyz_host.numel() * 4 + mask_host.numel()
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (3-way bf16 operand split, fp32 accumulate and residual stream)",
            "data": "synthetic",
            "config": workload_config(B, world, args.workload),
            "ms_per_unet_step": unet_ms,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="samples per GPU per step (BASELINE config 2: 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="uncond", choices=sorted(WORKLOADS),
                    help="uncond = the headline (BASELINE config 2); cam2lidar = BASELINE config 3(A), reported beside it")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help="bf16 = headline tensor-core path; fp32 = precise operand-split mode (parity mode)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
