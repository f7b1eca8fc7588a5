"""Developer tool: throughput of the Chamfer nearest-neighbour kernel (point pairs per second, CUDA events) next to the
numpy oracle on a bounded sample.  python tests/chamfer_bench.py [B] [N] [M]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from lidar_layout_b200.eval_ops import chamfer_3DDist
from oracle import eval_ref as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
M = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
a = torch.randn(B, N, 3, device="cuda") * 30
b = torch.randn(B, M, 3, device="cuda") * 30
mod = chamfer_3DDist()
for _ in range(3):
    mod(a, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    mod(a, b)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
pairs = 2.0 * B * N * M                      # both directions
print(f"lidm_chamfer_nn B={B} N={N} M={M}: {ms:.3f} ms per call, {pairs / ms / 1e9:.2f} T point pairs/s "
      f"(11 fp32-pipe instructions per pair and query: {11 * pairs / ms / 1e9:.1f} T lane-instructions/s of "
      f"~36 T issue peak = {11 * pairs / ms / 1e9 / 36:.2f})")
n = 4096
x, y = a[:1, :n].cpu().numpy(), b[:1, :n].cpu().numpy()
t0 = time.perf_counter()
E.chamfer_forward(x, y)
dt = time.perf_counter() - t0
print(f"numpy oracle, 1 core, {n}x{n} sample: {2.0 * n * n / dt / 1e9:.4f} G point pairs/s")
