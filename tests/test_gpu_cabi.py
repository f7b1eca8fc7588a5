"""GPU (-m gpu): error behaviour of the C ABI itself (include/lidm_b200.h) through raw ctypes: every failure is a
negative return code plus a message from lidm_last_error, never an exception across the boundary or a silent no-op."""
import ctypes
from ctypes import c_int64, c_void_p

import pytest
import torch

pytestmark = pytest.mark.gpu

from lidar_layout_b200 import _lib, config as C
from lidar_layout_b200.engine import to_cconfig
from lidar_layout_b200.weights import random_state_dict

ERR_INVALID, ERR_CUDA, ERR_STATE = -1, -2, -3


def _create(cfg):
    lib = _lib.load()
    h = c_void_p()
    cc = to_cconfig(cfg)
    return lib, h, lib.lidm_create(ctypes.byref(cc), ctypes.byref(h))


def _feed(lib, h, sd, skip=()):
    for name, t in sd.items():
        if name in skip:
            continue
        t = t.float().contiguous()
        shape = (c_int64 * max(t.dim(), 1))(*t.shape)
        assert lib.lidm_load_weight(h, name.encode(), c_void_p(t.data_ptr()), t.dim(), shape) == 0


def test_create_rejects_unsupported_configs(built_lib):
    lib = _lib.load()
    cfg = C.tiny()
    cc = to_cconfig(cfg)
    cc.num_head_channels = 64                      # the attention kernels are head-dim-32 only
    h = c_void_p()
    assert lib.lidm_create(ctypes.byref(cc), ctypes.byref(h)) == ERR_INVALID
    assert b"num_head_channels" in lib.lidm_last_error(None)
    cc = to_cconfig(cfg)
    cc.latent_w = 48                               # coarsest level would not tile into 128-pixel patches
    assert lib.lidm_create(ctypes.byref(cc), ctypes.byref(h)) == ERR_INVALID
    assert lib.lidm_create(None, ctypes.byref(h)) == ERR_INVALID


def test_call_order_and_missing_weights(built_lib):
    cfg = C.tiny()
    lib, h, rc = _create(cfg)
    assert rc == 0
    x = torch.zeros(1, *cfg.latent_shape, device="cuda")
    t = torch.zeros(1, dtype=torch.long, device="cuda")
    out = torch.empty_like(x)
    # forward before finalize
    assert lib.lidm_unet_forward(h, x.data_ptr(), t.data_ptr(), out.data_ptr(), 1, None) == ERR_STATE
    assert b"finalize" in lib.lidm_last_error(h)
    # finalize with a tensor missing: the message names it
    sd = random_state_dict(cfg, 0)
    missing = "model.diffusion_model.out.2.weight"
    _feed(lib, h, sd, skip=(missing,))
    assert lib.lidm_finalize_weights(h, 0) == ERR_STATE
    assert missing.encode() in lib.lidm_last_error(h)
    lib.lidm_destroy(h)
    # a wrongly shaped tensor is rejected at packing time
    lib, h, rc = _create(cfg)
    bad = dict(sd)
    bad[missing] = torch.zeros(3, 3)
    _feed(lib, h, bad)
    assert lib.lidm_finalize_weights(h, 0) == ERR_STATE
    lib.lidm_destroy(h)


def test_argument_checks_after_finalize(built_lib):
    cfg = C.tiny()
    lib, h, rc = _create(cfg)
    sd = random_state_dict(cfg, 0)
    _feed(lib, h, sd)
    assert lib.lidm_finalize_weights(h, 0) == 0
    assert lib.lidm_finalize_weights(h, 0) == ERR_STATE                       # twice
    t0 = sd["model.diffusion_model.out.2.bias"].float().contiguous()
    shape = (c_int64 * 1)(*t0.shape)
    assert lib.lidm_load_weight(h, b"x", c_void_p(t0.data_ptr()), 1, shape) == ERR_STATE   # load after finalize
    x = torch.zeros(1, *cfg.latent_shape, device="cuda")
    t = torch.zeros(1, dtype=torch.long, device="cuda")
    out = torch.empty_like(x)
    assert lib.lidm_unet_forward(h, None, t.data_ptr(), out.data_ptr(), 1, None) == ERR_INVALID      # null tensor
    assert lib.lidm_unet_forward(h, x.data_ptr(), t.data_ptr(), out.data_ptr(), 0, None) == ERR_INVALID   # empty batch
    # an unconditional model handed a context / a concat tensor
    ctx = torch.zeros(1, 4, 64, device="cuda")
    assert lib.lidm_unet_forward_cond(h, x.data_ptr(), t.data_ptr(), None, ctx.data_ptr(), 4, out.data_ptr(), 1, None) == ERR_INVALID
    assert lib.lidm_unet_forward_cond(h, x.data_ptr(), t.data_ptr(), x.data_ptr(), None, 0, out.data_ptr(), 1, None) == ERR_INVALID
    # encoder entry point without encoder weights
    img = torch.zeros(1, 1, *cfg.dataset.size, device="cuda")
    z = torch.empty(1, *cfg.latent_shape, device="cuda")
    assert lib.lidm_vq_encode(h, img.data_ptr(), z.data_ptr(), 1, None) == ERR_STATE
    # a good call still works afterwards (errors leave the handle usable)
    assert lib.lidm_unet_forward(h, x.data_ptr(), t.data_ptr(), out.data_ptr(), 1, None) == 0
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all())
    lib.lidm_destroy(h)


def test_stateless_entry_points_validate(built_lib):
    lib = _lib.load()
    assert lib.lidm_ddim_step(None, None, None, 0.5, 0.6, 0.0, 0.7, 1.0, None, None, 16, None) == ERR_INVALID
    img = torch.zeros(2, 8, 30, device="cuda")                                  # W not a multiple of 4
    xyz = torch.empty(2, 3, 8, 30, device="cuda")
    assert lib.lidm_backproject(img.data_ptr(), 2, 8, 30, 3.0, -25.0, 1.0, 56.0, 5.84, 1, 0, xyz.data_ptr(), None, None) == ERR_INVALID
    assert lib.lidm_compact_points(None, None, 1, 128, None, None, None) == ERR_INVALID
    assert lib.lidm_cfg_combine(None, 2.0, None, 16, None) == ERR_INVALID
