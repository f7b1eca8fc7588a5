"""lidar_layout_b200 — B200-native LiDM sampling path (DDIM loop -> VQ decode -> back-projection).

A drop-in for one hot path of AlanLiangC/LiDAR-Layout behind its own Python API:
    reference                                             here
    lidm.models.diffusion.ddpm.LatentDiffusion         -> lidar_layout_b200.LatentDiffusion
    lidm.models.diffusion.ddim.DDIMSampler             -> lidar_layout_b200.DDIMSampler
    lidm.utils.lidar_utils.range2pcd / range2xyz       -> lidar_layout_b200.range2pcd / range2xyz
    scripts/sample.py custom_to_pil / custom_to_pcd / save_logs -> lidar_layout_b200.postprocess.*
    lidm/eval/modules/chamfer{3D,2D} chamfer_3DDist / chamfer_2DDist (forward) -> lidar_layout_b200.eval_ops.*
All GPU work goes through the C ABI in include/lidm_b200.h (liblidm_b200.so, hand-written sm_100a kernels).
Importing the package never touches the GPU; using it without the built library or without a B200 raises.
"""
from . import config, schedule, weights          # noqa: F401  (host-only modules)
from ._lib import LidmError                      # noqa: F401


def __getattr__(name):
    # GPU-facing symbols are resolved lazily so that host-only tooling can import the package on CPU boxes.
    if name in ("LatentDiffusion",):
        from .ddpm import LatentDiffusion
        return LatentDiffusion
    if name in ("R2DMDiffusion",):
        from .ddpm import R2DMDiffusion
        return R2DMDiffusion
    if name in ("DDIMSampler",):
        from .ddim import DDIMSampler
        return DDIMSampler
    if name in ("range2pcd", "range2xyz", "range2xyz_gpu"):
        from . import lidar_utils
        return getattr(lidar_utils, name)
    if name in ("Engine",):
        from .engine import Engine
        return Engine
    if name in ("ops", "engine", "ddim", "ddpm", "lidar_utils", "postprocess", "parallel", "eval_ops"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
