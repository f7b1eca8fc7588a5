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


_N = {"bf16": "bf16", "fp16": "fp16", "fp32": "bf16x3 (3-way bf16 operand split, fp32 residual stream)"}
DTYPE_NAMES = {(u, a): (_N[u] if u == a else f"{_N[u]} U-Net + {_N[a]} first stage") + ", fp32 accumulate"
               for u in _N for a in _N}


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
    # BASELINE config 3(B): the layout-conditioned denoiser (LayoutDiffusionUNetModel + ObjectAwareCrossAttention), 32-beam
    # nuScenes range images (8x128 latents), synthetic LayoutTransformerEncoder outputs for 13 layout tokens
    "layout2lidar": ("nuscenes_layout2lidar", "layout-conditioned LiDM nuScenes 32x1024 (models/lidm/nuscenes/layout2lidar, "
                                              "LayoutDiffusionUNetModel, synthetic layout_outputs for 13 objects, random-init), "
                                              "DDIM-50 eta=0 + VQ decode + back-projection", 50 * 83.2 + 0.27 + 59.5),
    # BASELINE config 5: the R2DM pixel-space model (EfficientUNet, 2-channel 64x1024 range images), DDIM-256, no first stage
    "r2dm": ("nuscenes_r2dm", "R2DM pixel-space range-image diffusion (configs/r2dm, EfficientUNet, 64x1024, random-init), "
                              "DDIM-256 eta=0 + back-projection", 256 * 228.975),
}
WORKLOAD_STEPS = {"r2dm": 256}


def workload_config(B, world, name="uncond"):
    return {"workload": WORKLOADS[name][1],
            "batch_per_gpu": B, "global_batch": B * world, "ddim_steps": WORKLOAD_STEPS.get(name, DDIM_STEPS),
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


REF_BATCH = 2      # samples per CPU step: one step = REF_BATCH samples through all 50 DDIM steps + decode + back-projection


def cpu_step_seconds(cfg, sd, B, threads, n_unet_steps=DDIM_STEPS):
    t_unet, t_dec, t_bp = cpu_reference_run(cfg, sd, n_unet_steps, B, threads)
    return t_unet + t_dec + t_bp, dict(t_unet_per_step=t_unet / n_unet_steps, t_decode=t_dec, t_backproject=t_bp)


def run_reference_arm(args):
    """The reference algorithm on the host CPU (oracle port), every host thread.  One step = REF_BATCH samples through
    the WHOLE path (all 50 DDIM steps, decode, back-projection): a bounded sample of the B = 64 workload (the metric is
    per sample, and the CPU rate does not grow with the batch), timed as declared - no extrapolation."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from lidar_layout_b200 import config as C
    from lidar_layout_b200.weights import random_state_dict
    cfg = C.kitti_uncond()
    sd = random_state_dict(cfg, 0)
    threads = os.cpu_count() or 1
    B = REF_BATCH
    warm = min(args.warmup, 1)
    steps = max(1, min(args.steps, 3))        # ~20 s of host work per step on 16 cores
    for _ in range(warm):
        cpu_step_seconds(cfg, sd, B, threads, n_unet_steps=2)   # warm-up: page the weights in, size the thread pool
    t0 = time.perf_counter()
    for _ in range(steps):
        _, detail = cpu_step_seconds(cfg, sd, B, threads)
    sec = (time.perf_counter() - t0) / steps
    v = B / sec
    sample = (f"B={B} samples per step through all {DDIM_STEPS} DDIM steps + decode + back-projection (a bounded sample of the "
              f"B={args.batch} workload: samples/s is per sample); oracle port (torch fp32 CPU), {steps} timed steps, "
              f"per U-Net step {detail['t_unet_per_step']:.2f}s, decode {detail['t_decode']:.2f}s")
    conf = workload_config(B, 1)
    conf["sample_of"] = {"batch_per_gpu": args.batch, "n_gpus": args.gpus}
    conf["parallelism"] = f"host CPU, {threads} threads (one process whatever --gpus says)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1000.0 * sec, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": conf,
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ eager GPU baseline
def gpu_eager_baseline(cfg, sd, B, dev, unet_steps=2):
    """north_star target 1 ("reference-PyTorch-on-B200"): the reference algorithm as plain eager PyTorch on this GPU
    (the oracle's torch restatement moved to cuda: cuDNN / cuBLAS kernels, the reference's own op sequence), outside the
    timed region of the headline.  Bounded sample: `unet_steps` U-Net evaluations + DDIM updates and one decode at
    batch B per numeric setting; samples/s = B / (50 * t_unet_step + t_decode)."""
    from oracle import torch_ref as R
    sdd = {k: v.to(dev) for k, v in sd.items()}
    g = torch.Generator().manual_seed(1000)
    x0 = torch.randn((B,) + tuple(cfg.latent_shape), generator=g).to(dev)
    ts, table = R.ddim_schedule(cfg, DDIM_STEPS, 0.0)
    n = len(ts)
    out = {"batch": B, "sample": f"{unet_steps} U-Net evaluations + DDIM updates and one first-stage decode at B={B}, "
                                 f"extrapolated to {DDIM_STEPS} steps; oracle/torch_ref.py on cuda (eager cuDNN/cuBLAS)", "unit": "samples/s"}

    def one(name, tf32, autocast):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast else torch.autocast("cuda", enabled=False)
        try:
            with ctx, torch.no_grad():
                def step(x, i):
                    t = torch.full((B,), int(np.flip(ts)[i]), dtype=torch.long, device=dev)
                    e = R.unet_forward(sdd, cfg.unet, x, t).float()
                    return R.ddim_step(x, e, table[n - 1 - i])[0]
                x = step(x0, 0)                     # warm-up (cuDNN autotune, allocator)
                R.decode_first_stage(sdd, cfg, x0)
                torch.cuda.synchronize()
                a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                a.record()
                x = x0
                for i in range(unet_steps):
                    x = step(x, i)
                b.record()
                R.decode_first_stage(sdd, cfg, x)
                c.record()
                torch.cuda.synchronize()
            t_step, t_dec = a.elapsed_time(b) / unet_steps, b.elapsed_time(c)
            out[name] = {"value": B / ((DDIM_STEPS * t_step + t_dec) / 1e3), "ms_per_unet_step": t_step, "decode_ms": t_dec}
        except Exception as ex:                     # e.g. out of memory at a large batch: report, do not fail the bench
            out[name] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}
        torch.cuda.empty_cache()

    one("fp32", False, False)
    one("tf32", True, False)
    one("autocast_bf16", True, True)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = True
    return out


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

    r2dm = args.workload == "r2dm"
    cfg = C.nuscenes_r2dm((64, 1024)) if r2dm else getattr(C, WORKLOADS[args.workload][0])()
    B = args.batch
    n_ddim = WORKLOAD_STEPS.get(args.workload, DDIM_STEPS)
    sd = random_state_dict(cfg, 0)
    if r2dm:
        model = L.R2DMDiffusion(cfg, device=dev, use_ema=False, precision=None if args.precision == "bf16" else args.precision)
    else:
        model = L.LatentDiffusion(cfg, device=dev, use_ema=False, precision=args.precision, ae_precision=args.ae_precision)
    model.load_state_dict(sd)
    cfg = model.cfg
    sampler = L.DDIMSampler(model)
    sampler.make_schedule(n_ddim, ddim_eta=0.0)
    ts, table = sampler.ddim_timesteps, sampler.ddim_table
    assert len(ts) == n_ddim
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
    elif cfg.conditioning_key == "layout_crossattn":
        # what LayoutTransformerEncoder.forward returns (layout_encoder.py:239-279) for 13 layout tokens, synthetic
        gctx = torch.Generator().manual_seed(1007)
        E, u = cfg.unet.encoder_channels, cfg.unet
        mk = lambda *shape: parallel.local_slice(torch.randn((global_B,) + shape, generator=gctx), rank, world).to(dev)
        cond = {"xf_proj": 0.1 * mk(u.time_embed_dim), "xf_out": mk(E, 13), "obj_class_embedding": mk(E, 13),
                "obj_bbox_embedding": mk(E, 13)}
        for r in (4, 2, 1):
            cond[f"image_patch_bbox_embedding_for_resolution{r}"] = torch.randn((1, E, r * 16 * r), generator=gctx).to(dev)
        cond_kw = dict(layout_cond=cond)

    def decode(z):      # first-stage decode; the pixel-space model's sample is the image: channel 0 = depth
        return z[:, :1].contiguous() if r2dm else eng.vq_decode(z)

    def device_step():
        z, _ = eng.ddim_sample(x_T_dev, ts, table, **cond_kw)
        img = decode(z)
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
            z, _ = sampler.sample(n_ddim, batch_size=B, shape=cfg.latent_shape, eta=0.0, x_T=x, conditioning=cond)
        img = z[:, :1].contiguous() if r2dm else model.decode_first_stage(z)
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
        img = decode(z)
        prof_dec = _lib.profile_end()
        peaks = measured_peaks()
        g = prof_unet["conv_gemm"]
        ach = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        unet_ms = sum(v["ms"] for v in prof_unet.values()) / n_ddim
        roof = {
            "bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM conv, all U-Net launches of one DDIM-50 loop)",
            "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": ach / peaks["bf16_tflops_sustained"], "peak_source": peaks["source"] + " (sustained: kernel timed inside a long step)",
            "traffic": captured_traffic(),
            "traffic_unit": "bytes per launch (mean of the launches captured in profiles/roofline_traffic.json)",
            "launches": g["launches"], "avg_launch_ms": g["ms"] / max(g["launches"], 1),
            "share_of_step": {k: v["ms"] / max(sum(x["ms"] for x in prof_unet.values()), 1e-9) for k, v in prof_unet.items()},
            "groupnorm_GBps": (prof_unet["groupnorm"]["bytes"] / (prof_unet["groupnorm"]["ms"] * 1e-3) / 1e9
                               if prof_unet["groupnorm"]["ms"] > 0 else None),
            "attention_TFLOPs": (prof_unet["attention"]["flops"] / (prof_unet["attention"]["ms"] * 1e-3) / 1e12
                                 if prof_unet["attention"]["ms"] > 0 else None),
            "decoder_ms": sum(v["ms"] for v in prof_dec.values()),
            "whole_pipeline_frac_of_tensor_peak": value / world * WORKLOADS[args.workload][2] / 1e3 / peaks["bf16_tflops_sustained"],
        }

    # strong scaling (BASELINE config 2 is a GLOBAL batch of 64 on 8 GPUs): the same global batch as one GPU's step, split
    # over the ranks.  speedup_vs_n1 is against this run's own per-GPU rate (each rank's weak step IS the N = 1 workload).
    strong = None
    if world > 1 and B % world == 0:
        Bs = B // world
        xs_dev = parallel.local_slice(x_T_global[:B], rank, world).to(dev)
        cs_kw = dict(context=parallel.local_slice(ctx_global[:B], rank, world).to(dev)) if cfg.conditioning_key == "crossattn" else {}
        if cfg.conditioning_key == "layout_crossattn":
            strong = None
            cs_kw = None

        def strong_step():
            z, _ = eng.ddim_sample(xs_dev, ts, table, **cs_kw)
            img = decode(z)
            xyz, mask = L.ops.backproject(img, ds.fov, ds.depth_range, ds.depth_scale, ds.log_scale)
            return parallel.all_gather_batch(img, B), xyz, mask

        if cs_kw is not None:
            for _ in range(2):
                strong_step()
            ms_s = timed(strong_step, args.steps)
            sv = B * args.steps / (ms_s / 1000.0)
            strong = {"global_batch": B, "batch_per_gpu": Bs, "value": sv, "unit": "samples/s", "ms_per_step": ms_s / args.steps,
                      "speedup_vs_n1": sv / (value / world)}

    # the other numeric modes, same step, fewer repeats (outside the headline's timed region)
    modes = None
    if rank == 0 and world == 1 and not args.no_modes and args.workload == "uncond":
        modes = {}
        for key, (pu, pa) in {"precise": ("fp32", "fp32"), "fp16": ("fp16", "fp16")}.items():
            if (pu, pa) == (cfg.precision, cfg.ae_precision_resolved):
                continue
            m2 = L.LatentDiffusion(cfg, device=dev, use_ema=False, precision=pu, ae_precision=pa)
            m2.load_state_dict(sd)
            e2 = m2.engine

            def mode_step():
                z, _ = e2.ddim_sample(x_T_dev, ts, table, **cond_kw)
                img = e2.vq_decode(z)
                return L.ops.backproject(img, ds.fov, ds.depth_range, ds.depth_scale, ds.log_scale)

            mode_step()
            ms_m = timed(mode_step, 2)
            modes[key] = {"value": B * 2 / (ms_m / 1000.0), "unit": "samples/s", "ms_per_step": ms_m / 2,
                          "dtype": DTYPE_NAMES[(pu, pa)]}
            del m2, e2
            torch.cuda.empty_cache()

    eager = None
    if rank == 0 and world == 1 and not args.no_eager_baseline and args.workload == "uncond":
        eager = gpu_eager_baseline(cfg, sd, B, dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "uncond":
        threads = os.cpu_count() or 1
        sec, detail = cpu_step_seconds(cfg, sd, REF_BATCH, threads, n_unet_steps=10)
        full = detail["t_unet_per_step"] * DDIM_STEPS + detail["t_decode"] + detail["t_backproject"]
        cpu = {"value": REF_BATCH / full, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"B={REF_BATCH}, 10 of {DDIM_STEPS} DDIM steps timed and extrapolated x5, full decode + back-projection "
                         f"(oracle port, torch fp32); per-step {detail['t_unet_per_step']:.2f}s decode {detail['t_decode']:.2f}s "
                         f"(`--impl reference` times all {DDIM_STEPS} steps)"}

    if rank == 0:
        h2d = x_T_host.numel() * 4
        d2h = xyz_host.numel() * 4 + mask_host.numel()
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": (cfg.precision + ", fp32 accumulate") if r2dm else DTYPE_NAMES[(cfg.precision, cfg.ae_precision_resolved)],
            "data": "synthetic",
            "config": workload_config(B, world, args.workload),
            "ms_per_unet_step": unet_ms,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roof,
            "cpu_baseline": cpu,
            "gpu_eager_baseline": eager,
            "precise": modes.get("precise") if modes else None,
            "modes": modes,
            "strong_scaling": strong,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_ae_arm(args):
    """BASELINE config 4: first-stage autoencoder throughput - VQModel encode (Encoder -> quant_conv) + decode
    (quantise -> post_quant_conv -> Decoder) of 64x1024 range images, `--batch` images per GPU and step (default 64), inputs
    clip(N(0, 0.5), -1, 1).  Same line format as the headline: `value` with the images resident in HBM, `e2e` through the
    reference-facing calls (encode_first_stage / decode_first_stage) with pinned host buffers in and out."""
    import torch
    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py --workload ae needs a B200"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import lidar_layout_b200 as L
    from lidar_layout_b200 import _lib, config as C
    from lidar_layout_b200.weights import random_encoder_state_dict, random_state_dict
    cfg = C.kitti_uncond()
    B = args.batch
    model = L.LatentDiffusion(cfg, device=dev, use_ema=False, precision=args.precision, ae_precision=args.ae_precision)
    sd = {**random_state_dict(cfg, 0), **random_encoder_state_dict(cfg, 0)}
    model.load_state_dict(sd)
    cfg = model.cfg
    eng = model.engine
    H, W = cfg.dataset.size
    g = torch.Generator().manual_seed(1000 + rank)
    x_host = (torch.randn((B, 1, H, W), generator=g) * 0.5).clamp_(-1, 1).pin_memory()
    x_dev = x_host.to(dev)
    out_host = torch.empty((B, 1, H, W), dtype=torch.float32).pin_memory()

    def device_step():
        return eng.vq_decode(eng.vq_encode(x_dev))

    def e2e_step():
        x = x_host.to(dev, non_blocking=True)
        rec = model.decode_first_stage(model.encode_first_stage(x))
        out_host.copy_(rec, non_blocking=True)
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

    steps = max(args.steps, 1) * 10          # one step is ~40 ms: time ten per requested step
    for _ in range(max(args.warmup, 3)):
        device_step()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    ms = timed(device_step, steps)
    launches = _lib.launch_count() - l0
    clk = clocks.stop() if rank == 0 else {}
    value = B * world * steps / (ms / 1000.0)
    e2e_step()
    ms_e2e = timed(e2e_step, steps)
    roof = cpu = None
    if rank == 0:
        torch.cuda.synchronize()
        _lib.profile_begin()
        z = eng.vq_encode(x_dev)
        prof_enc = _lib.profile_end()
        _lib.profile_begin()
        eng.vq_decode(z)
        prof_dec = _lib.profile_end()
        peaks = measured_peaks()
        gm = {k: prof_enc["conv_gemm"][k] + prof_dec["conv_gemm"][k] for k in ("flops", "ms", "launches")}
        gn = {k: prof_enc["groupnorm"][k] + prof_dec["groupnorm"][k] for k in ("bytes", "ms")}
        ach = gm["flops"] / (gm["ms"] * 1e-3) / 1e12 if gm["ms"] > 0 else 0.0
        tot = sum(v["ms"] for v in prof_enc.values()) + sum(v["ms"] for v in prof_dec.values())
        roof = {"bound": "tensor", "kernel": "conv_gemm_kernel (all encoder + decoder launches of one step)", "achieved": ach,
                "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"],
                "peak_source": peaks["source"] + " (sustained)", "traffic": None, "launches": gm["launches"],
                "avg_launch_ms": gm["ms"] / max(gm["launches"], 1),
                "share_of_step": {"conv_gemm": gm["ms"] / tot, "groupnorm": gn["ms"] / tot},
                "groupnorm_GBps": gn["bytes"] / (gn["ms"] * 1e-3) / 1e9 if gn["ms"] > 0 else None,
                "encode_ms": sum(v["ms"] for v in prof_enc.values()), "decode_ms": sum(v["ms"] for v in prof_dec.values()),
                "note": "92.6 + 119.1 GFLOP per image (SURVEY section 8d); the 64 / 128-channel full-resolution levels are HBM-bound"}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import torch_ref as R
            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            xs = x_host[:1].clone()
            with torch.no_grad():
                t0 = time.perf_counter()
                zz = R.encode_first_stage(sd, cfg, xs)
                R.decode_first_stage(sd, cfg, zz)
                dt = time.perf_counter() - t0
            cpu = {"value": 1.0 / dt, "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": "B=1: one encode + quantised decode (oracle port, torch fp32)"}
        line = {
            "metric": "images/sec first-stage VQ autoencoder 64x1024 (encode + quantised decode)", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_NAMES[(cfg.precision, cfg.ae_precision_resolved)].split(" U-Net + ")[-1],
            "data": "synthetic",
            "config": {"workload": "first-stage autoencoder f_c2_p4 (models/first_stage_models/kitti/f_c2_p4, random-init), "
                                   "VQModel encode + decode of 64x1024 range images (BASELINE config 4)",
                       "batch_per_gpu": B, "global_batch": B * world,
                       "l2": "activations of one step (GBs) >> 126 MB L2", "parallelism": f"batch-sharded x{world}, no collective"},
            "e2e": {"value": B * world * steps / (ms_e2e / 1000.0), "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e2e / steps},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
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
    ap.add_argument("--workload", default="uncond", choices=sorted(WORKLOADS) + ["ae"],
                    help="uncond = the headline (BASELINE config 2); cam2lidar = BASELINE config 3(A), reported beside it")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "fp16"],
                    help="U-Net numeric mode: bf16 = headline tensor-core path; fp32 = precise operand-split mode; fp16 = IEEE half")
    ap.add_argument("--ae-precision", default=None, choices=["bf16", "fp32", "fp16"],
                    help="first-stage numeric mode (default: fp16 under a bf16 U-Net, else the U-Net's)")
    ap.add_argument("--no-modes", action="store_true", help="skip the precise / fp16 mode timings")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the eager-PyTorch-on-GPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "ae":
        run_ae_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
