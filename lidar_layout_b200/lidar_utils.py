"""range2pcd / range2xyz with the reference's signatures (lidm/utils/lidar_utils.py:134-204), computed by the
back-projection kernel (lidm_backproject).  NumPy in / NumPy out like the reference; `*_gpu` variants keep
tensors on the device for whole batches."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _to_dev(range_img):
    if isinstance(range_img, torch.Tensor):
        t = range_img
    else:
        t = torch.from_numpy(np.ascontiguousarray(range_img, dtype=np.float32))
    if not t.is_cuda:
        t = t.cuda()
    return t.float()


def range2xyz_gpu(range_img: torch.Tensor, fov, depth_range, depth_scale, log_scale=True, input_is_unit=True):
    """(B,H,W) or (H,W) device tensor -> ((B,)3,H,W) fp32 xyz with -1 where masked, plus uint8 mask."""
    t = _to_dev(range_img)
    squeeze = t.dim() == 2
    if squeeze:
        t = t[None]
    xyz, mask = ops.backproject(t, fov, depth_range, depth_scale, log_scale, True, input_is_unit=input_is_unit)
    return (xyz[0], mask[0]) if squeeze else (xyz, mask)


def range2pcd(range_img, fov, depth_range, depth_scale, log_scale=True, label=None, color=None, **kwargs):
    """reference lidar_utils.py:134-172: range_img (H,W) in [0,1] -> (pcd (N,3) float64, color, label)."""
    xyz, mask = range2xyz_gpu(range_img, fov, depth_range, depth_scale, log_scale)
    m = mask.flatten().bool()
    pcd = xyz.reshape(3, -1).t()[m].double().cpu().numpy()
    mk = m.cpu().numpy()
    if label is not None:
        label = np.asarray(label).flatten()[mk]
    if color is not None:
        color = np.asarray(color).reshape(-1, 3)[mk, :]
    else:
        color = np.ones((pcd.shape[0], 3)) * [0.7, 0.7, 1]
    return pcd, color, label


def range2xyz(range_img, fov, depth_range, depth_scale, log_scale=True, **kwargs):
    """reference lidar_utils.py:175-204: (H,W) in [0,1] -> (3,H,W) float64, -1 where masked."""
    xyz, _ = range2xyz_gpu(range_img, fov, depth_range, depth_scale, log_scale)
    return xyz.double().cpu().numpy()
