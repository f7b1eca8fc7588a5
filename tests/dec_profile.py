"""Developer tool: per-op CUDA-event breakdown of the first-stage decoder (and encoder) at batch B."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_layout_b200 import _lib, config as C
from lidar_layout_b200.engine import Engine
from lidar_layout_b200.weights import random_encoder_state_dict, random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
which = sys.argv[2] if len(sys.argv) > 2 else "decode"
path = os.environ.setdefault("LIDM_PROFILE_DUMP", "/tmp/dec_ops.csv")
if os.path.exists(path):
    os.unlink(path)
cfg = C.kitti_uncond()
eng = Engine(cfg).load_state_dict({**random_state_dict(cfg, 0), **random_encoder_state_dict(cfg, 0)})
z = torch.randn(B, 8, 16, 128, device="cuda")
img = torch.randn(B, 1, 64, 1024, device="cuda").clamp_(-1, 1)
fn = (lambda: eng.vq_decode(z)) if which == "decode" else (lambda: eng.vq_encode(img))
for _ in range(2):
    fn()
torch.cuda.synchronize()
_lib.profile_begin()
for _ in range(2):
    fn()
_lib.profile_end()
agg = collections.OrderedDict()
for line in open(path):
    cat, ms, fl, by, label = line.rstrip("\n").split(",", 4)
    a = agg.setdefault(label or ("other cat" + cat), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += float(ms); a[2] += float(fl); a[3] += float(by)
tot = sum(a[1] for a in agg.values())
print(f"{which} B={B}: {tot/2:.3f} ms per call")
for label, (n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    rate = f"{fl/ms/1e9:8.1f} TF/s" if fl > 0 else (f"{by/ms/1e6:8.1f} GB/s" if by > 0 else " " * 13)
    print(f"{ms/2:8.3f} ms {100*ms/tot:5.1f}%  n={n//2:3d}  avg {1000*ms/n:8.1f} us  {rate}  {label}")
