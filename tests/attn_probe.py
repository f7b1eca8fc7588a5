"""Developer probe: the fused attention kernel alone at the U-Net's level-0 shape (for ncu captures / timing).
    python tests/attn_probe.py [B] [heads] [T] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_layout_b200 import ops

B, heads, T, reps = (int(a) for a in (sys.argv[1:5] + ["64", "8", "2048", "5"][len(sys.argv) - 1:]))
torch.manual_seed(0)
qkv = torch.randn(B, heads * 96, T, device="cuda")
for _ in range(2):
    y = ops.qkv_attention_legacy(qkv, heads)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    y = ops.qkv_attention_legacy(qkv, heads)
b.record()
torch.cuda.synchronize()
print(f"attention op (incl. layout conversions) B{B} heads{heads} T{T}: {a.elapsed_time(b) / reps:.3f} ms per call")
