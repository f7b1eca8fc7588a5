"""Sample post-processing with the reference script's interface (scripts/sample.py:29-56, 141-162): the step that
immediately follows the sampling path.  The per-image work (uint8 conversion, back-projection, `pcd[mask]` gather) runs
for the whole batch on the device (lidm_to_uint8_image, lidm_backproject, lidm_compact_points); only the finished
arrays cross PCIe, once per batch, and the on-disk formats are the reference's own (PNG via PIL, `np.savetxt('%.3f')`
xyz+rgb text, `joblib.dump` of a list of float32 (N_i, 3) arrays)."""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np
import torch

from . import ops


def _dataset_kwargs(config):
    """`config['data']['params']['dataset']` of the reference YAML, a LidmConfig, or a DatasetConfig."""
    ds = config
    if hasattr(config, "dataset"):
        ds = config.dataset
    elif isinstance(config, dict) or hasattr(config, "keys"):
        ds = config["data"]["params"]["dataset"] if "data" in config else config
    get = (lambda k: ds[k]) if (isinstance(ds, dict) or hasattr(ds, "keys")) else (lambda k: getattr(ds, k))
    return dict(fov=tuple(get("fov")), depth_range=tuple(get("depth_range")), depth_scale=float(get("depth_scale")),
                log_scale=bool(get("log_scale")))


def _as_batch(x: torch.Tensor) -> torch.Tensor:
    """(H,W) / (1,H,W) / (B,1,H,W) / (B,H,W) -> (B,H,W) CUDA fp32."""
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.asarray(x))
    if not x.is_cuda:
        x = x.cuda()
    x = x.detach().float()
    if x.dim() == 4:
        assert x.shape[1] == 1, "range images have one channel"
        x = x[:, 0]
    elif x.dim() == 2:
        x = x[None]
    elif x.dim() == 3 and x.shape[0] == 1:
        pass
    return x.contiguous()


def custom_to_np(x):
    """scripts/sample.py:48-52."""
    x = x.detach().cpu().squeeze().numpy()
    return (np.clip(x, -1., 1.) + 1.) / 2.


def images_to_uint8(batch) -> np.ndarray:
    """custom_to_pil's array for a whole batch: (B,H,W) uint8 on the host (one device pass, one copy)."""
    return ops.to_uint8_image(_as_batch(batch)).cpu().numpy()


def custom_to_pil(x):
    """scripts/sample.py:38-45."""
    from PIL import Image
    return Image.fromarray(images_to_uint8(x)[0])


def samples_to_point_clouds(batch, config) -> List[np.ndarray]:
    """[custom_to_pcd(img, config)[0].astype(np.float32) for img in batch] (scripts/sample.py:131) for the whole
    batch: back-projection and the valid-point gather run on the device, one (N_i, 3) float32 array per sample."""
    kw = _dataset_kwargs(config)
    x = _as_batch(batch)
    xyz, mask = ops.backproject(x, kw["fov"], kw["depth_range"], kw["depth_scale"], kw["log_scale"], return_mask=True)
    points, counts = ops.compact_points(xyz, mask)
    counts_h = counts.cpu().numpy()
    nmax = int(counts_h.max()) if counts_h.size else 0
    pts_h = points[:, :nmax].cpu().numpy() if nmax else np.zeros((x.shape[0], 0, 3), np.float32)
    return [np.ascontiguousarray(pts_h[b, :int(counts_h[b])]) for b in range(x.shape[0])]


def custom_to_pcd(x, config):
    """scripts/sample.py:29-35: one range image -> (xyz (N,3) float64, rgb zeros)."""
    xyz = samples_to_point_clouds(x, config)[0].astype(np.float64)
    return xyz, np.zeros_like(xyz)


def save_logs(logs, imglogdir, pcdlogdir, n_saved=0, key="samples", np_path=None, config=None):
    """scripts/sample.py:141-162, same files and formats; the batch is processed on the device in one pass."""
    for k in logs:
        if k != key:
            continue
        batch = logs[key]
        if np_path is None:
            from PIL import Image
            imgs = images_to_uint8(batch)
            clouds = samples_to_point_clouds(batch, config)
            for img, xyz in zip(imgs, clouds):
                Image.fromarray(img).save(os.path.join(imglogdir, f"{key}_{n_saved:06}.png"))
                xyz = xyz.astype(np.float64)
                np.savetxt(os.path.join(pcdlogdir, f"{key}_{n_saved:06}.txt"), np.hstack([xyz, np.zeros_like(xyz)]),
                           fmt='%.3f')
                n_saved += 1
        else:
            npbatch = custom_to_np(batch)
            shape_str = "x".join([str(x) for x in npbatch.shape])
            np.savez(os.path.join(np_path, f"{n_saved}-{shape_str}-samples.npz"), npbatch)
            n_saved += npbatch.shape[0]
    return n_saved


def dump_point_clouds(all_samples: Sequence[np.ndarray], path: str):
    """joblib.dump(all_samples, '<nplog>/samples.pcd') (scripts/sample.py:135): the file `--eval` / `-f` consume."""
    import joblib
    joblib.dump(list(all_samples), path)
