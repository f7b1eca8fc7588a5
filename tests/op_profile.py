"""Developer tool: per-op CUDA-event breakdown of one U-Net evaluation (and optionally the decoder) at batch B.
    LIDM_PROFILE_DUMP=gpurun_out/ops.csv python tests/op_profile.py [B] [uncond|cam2lidar|sem2lidar|layout2lidar|r2dm] [ctx_len]"""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lidar_layout_b200 import _lib, config as C
from lidar_layout_b200.engine import Engine
from lidar_layout_b200.weights import random_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
path = os.environ.setdefault("LIDM_PROFILE_DUMP", "/tmp/ops.csv")
if os.path.exists(path):
    os.unlink(path)
which = sys.argv[2] if len(sys.argv) > 2 else "uncond"
L = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg = {"uncond": C.kitti_uncond, "cam2lidar": C.kitti_cam2lidar, "sem2lidar": C.kitti_sem2lidar,
       "layout2lidar": C.nuscenes_layout2lidar, "r2dm": lambda: C.nuscenes_r2dm((64, 1024))}[which]()
eng = Engine(cfg).load_state_dict(random_state_dict(cfg, 0))
x = torch.randn((B,) + tuple(cfg.latent_shape), device="cuda")
t = torch.full((B,), 501, dtype=torch.long, device="cuda")
kw = {}
if which == "cam2lidar":
    kw["context"] = torch.randn(B, L, cfg.unet.context_dim, device="cuda")
if which == "sem2lidar":
    kw["c_concat"] = torch.randn(B, cfg.unet.in_channels - 8, 16, 128, device="cuda")
if which == "layout2lidar":
    E = cfg.unet.encoder_channels
    cond = {"xf_proj": 0.1 * torch.randn(B, cfg.unet.time_embed_dim, device="cuda")}
    for k in ("xf_out", "obj_class_embedding", "obj_bbox_embedding"):
        cond[k] = torch.randn(B, E, 13, device="cuda")
    for r in (4, 2, 1):
        cond[f"image_patch_bbox_embedding_for_resolution{r}"] = torch.randn(1, E, 16 * r * r, device="cuda")
    kw["layout_cond"] = cond
for _ in range(3):
    eng.unet_forward(x, t, **kw)
torch.cuda.synchronize()
_lib.profile_begin()
for _ in range(3):
    eng.unet_forward(x, t, **kw)
res = _lib.profile_end()
agg = collections.OrderedDict()
for line in open(path):
    cat, ms, fl, by, label = line.rstrip("\n").split(",", 4)
    a = agg.setdefault(label or "other", [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += float(ms); a[2] += float(fl); a[3] += float(by)
tot = sum(a[1] for a in agg.values())
print(f"U-Net {which} B={B}: {tot/3:.3f} ms per evaluation (3 runs)")
for label, (n, ms, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    rate = f"{fl/ms/1e9:8.1f} TF/s" if fl > 0 else (f"{by/ms/1e6:8.1f} GB/s" if by > 0 else " " * 13)
    print(f"{ms/3:8.3f} ms {100*ms/tot:5.1f}%  n={n//3:3d}  avg {1000*ms/n:8.1f} us  {rate}  {label}")
