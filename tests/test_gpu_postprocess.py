"""GPU (-m gpu): sample post-processing against the oracle's restatement of the reference
script (scripts/sample.py:29-45, 129-162): uint8 images and the compaction order are bit-exact (integer / index work);
point coordinates within 2e-5 of the reference's float64 result (fp32 output, as for lidm_backproject)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from lidar_layout_b200 import config as C
from oracle import torch_ref as R


def _images(B, H, W, seed=0):
    rng = np.random.Generator(np.random.PCG64(seed))
    x = np.clip(rng.standard_normal((B, 1, H, W), dtype=np.float32) * 0.7, -1.3, 1.3).astype(np.float32)
    x[0, 0, :2] = -1.0          # empty rows
    if B > 1:
        x[1] = -1.0             # a sample without a single valid point
    return x


@pytest.mark.parametrize("shape", [(3, 64, 1024), (2, 16, 256), (1, 5, 12)])
def test_uint8_image_bit_exact(built_lib, shape):
    from lidar_layout_b200 import ops
    x = _images(*shape)
    got = ops.to_uint8_image(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = (255 * ((np.clip(x, -1., 1.) + 1.) / 2.)).astype(np.uint8)     # custom_to_pil, scripts/sample.py:38-45
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("shape", [(3, 64, 1024), (2, 16, 256)])
def test_point_clouds_match_reference_order_and_values(built_lib, shape):
    from lidar_layout_b200 import postprocess
    cfg = C.kitti_uncond()
    x = _images(*shape, seed=3)
    clouds = postprocess.samples_to_point_clouds(torch.from_numpy(x).cuda(), cfg)
    ds = dict(fov=cfg.dataset.fov, depth_range=cfg.dataset.depth_range, depth_scale=cfg.dataset.depth_scale,
              log_scale=cfg.dataset.log_scale)
    assert len(clouds) == shape[0]
    for b in range(shape[0]):
        ref, mask = R.range2pcd(R.custom_to_unit(x[b, 0]), **ds)
        assert clouds[b].dtype == np.float32 and clouds[b].shape == ref.shape       # same points, same count
        if ref.shape[0]:
            assert np.abs(clouds[b].astype(np.float64) - ref).max() < 2e-5            # same order (row-major)
    assert clouds[1].shape[0] == 0


def test_compaction_equals_boolean_indexing(built_lib):
    from lidar_layout_b200 import ops
    torch.manual_seed(0)
    B, H, W = 4, 64, 1024
    xyz = torch.randn(B, 3, H, W, device="cuda")
    mask = (torch.rand(B, H, W, device="cuda") < 0.37).to(torch.uint8)
    mask[2] = 1
    mask[3] = 0
    pts, cnt = ops.compact_points(xyz, mask)
    for b in range(B):
        ref = xyz[b].reshape(3, -1).t()[mask[b].flatten().bool()]
        assert int(cnt[b]) == ref.shape[0]
        assert torch.equal(pts[b, :ref.shape[0]], ref)          # bit-exact gather, numpy/torch mask order


def test_save_logs_writes_reference_formats(built_lib, tmp_path):
    from PIL import Image
    from lidar_layout_b200 import postprocess
    cfg = C.kitti_uncond()
    x = _images(2, 64, 1024, seed=5)
    imgdir, pcddir = tmp_path / "img", tmp_path / "pcd"
    imgdir.mkdir(); pcddir.mkdir()
    n = postprocess.save_logs({"samples": torch.from_numpy(x).cuda(), "other": None}, str(imgdir), str(pcddir), n_saved=7,
                              key="samples", config=cfg)
    assert n == 9
    ds = dict(fov=cfg.dataset.fov, depth_range=cfg.dataset.depth_range, depth_scale=cfg.dataset.depth_scale,
              log_scale=cfg.dataset.log_scale)
    for i in range(2):
        png = np.asarray(Image.open(imgdir / f"samples_{7 + i:06}.png"))
        assert np.array_equal(png, (255 * R.custom_to_unit(x[i, 0])).astype(np.uint8))
        ref, _ = R.range2pcd(R.custom_to_unit(x[i, 0]), **ds)
        txt = np.loadtxt(pcddir / f"samples_{7 + i:06}.txt", ndmin=2)
        if ref.shape[0] == 0:
            assert txt.size == 0
            continue
        assert txt.shape == (ref.shape[0], 6) and not txt[:, 3:].any()                # xyz + zero rgb, '%.3f'
        assert np.abs(txt[:, :3] - ref).max() <= 1.1e-3
    clouds = postprocess.samples_to_point_clouds(torch.from_numpy(x).cuda(), cfg)
    postprocess.dump_point_clouds(clouds, str(tmp_path / "samples.pcd"))
    import joblib
    back = joblib.load(tmp_path / "samples.pcd")
    assert len(back) == 2 and all(np.array_equal(a, b) for a, b in zip(back, clouds))
