"""Developer probe: fused attention kernel on a list of (B, heads, T) cases, each in its own subprocess."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def run(B, heads, T, scale):
    import torch
    from lidar_layout_b200 import ops
    from oracle import torch_ref as R
    torch.manual_seed(0)
    qkv = torch.randn(B, heads * 96, T) * scale
    y = ops.qkv_attention_legacy(qkv.cuda(), heads)
    torch.cuda.synchronize()
    ref = R.qkv_attention_legacy(qkv.bfloat16().float(), heads)
    d = (y.cpu().double() - ref.double()).flatten()
    print(f"attn B{B} heads{heads} T{T} scale{scale} rel {float(d.norm() / ref.double().norm()):.2e}", flush=True)

if __name__ == "__main__":
    if sys.argv[1] == "--run":
        run(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]))
        sys.exit(0)
    cases = [tuple(float(v) if "." in v else int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for c in cases:
        c = c + (1.0,) if len(c) == 3 else c
        r = subprocess.run([sys.executable, __file__, "--run"] + [str(v) for v in c], capture_output=True, text=True, timeout=120)
        out = (r.stdout + r.stderr).strip().splitlines()
        print(f"{c}: rc={r.returncode} " + (out[-1] if r.returncode == 0 else " | ".join(l for l in out if "timeout" in l or "Error" in l)[:600]), flush=True)
