"""ctypes binding of liblidm_b200.so (C ABI in include/lidm_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or a call fails the
error is raised to the caller.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblidm_b200.so")
MAX_LEVELS = 8


class LidmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"lidm_b200 error {code}: {msg}")
        self.code = code


class CConfig(ctypes.Structure):
    """struct lidm_config (include/lidm_b200.h)."""
    _fields_ = [
        ("in_channels", c_int32), ("out_channels", c_int32), ("model_channels", c_int32),
        ("num_res_blocks", c_int32), ("num_head_channels", c_int32),
        ("n_channel_mult", c_int32), ("channel_mult", c_int32 * MAX_LEVELS),
        ("n_attention_resolutions", c_int32), ("attention_resolutions", c_int32 * MAX_LEVELS),
        ("latent_h", c_int32), ("latent_w", c_int32),
        ("embed_dim", c_int32), ("n_embed", c_int32), ("z_channels", c_int32), ("ae_ch", c_int32),
        ("ae_out_ch", c_int32), ("ae_num_res_blocks", c_int32), ("ae_use_mask", c_int32),
        ("ae_n_ch_mult", c_int32), ("ae_ch_mult", c_int32 * MAX_LEVELS),
        ("ae_strides", (c_int32 * 2) * MAX_LEVELS),
        ("scale_factor", c_float),
        ("precision", c_int32),
        ("latent_channels", c_int32),
        ("use_spatial_transformer", c_int32), ("context_dim", c_int32), ("transformer_depth", c_int32),
        ("ae_in_channels", c_int32),
        ("ae_precision", c_int32),
        ("unet_type", c_int32), ("encoder_channels", c_int32), ("num_attention_blocks", c_int32),
        ("enc_layers", c_int32), ("enc_heads", c_int32), ("enc_out_dim", c_int32), ("enc_num_classes", c_int32),
        ("eff_res_blocks", c_int32 * MAX_LEVELS), ("eff_gn_groups", c_int32), ("eff_attn_heads", c_int32),
        ("eff_gn_eps", c_float),
    ]


EXPORTS = [
    "lidm_last_error", "lidm_create", "lidm_destroy", "lidm_load_weight", "lidm_finalize_weights",
    "lidm_unet_forward", "lidm_unet_forward_cond", "lidm_ddim_step", "lidm_ddpm_step", "lidm_ddim_sample", "lidm_ddim_sample_cond",
    "lidm_cfg_combine", "lidm_layout_set_cond", "lidm_layout_encode", "lidm_vq_decode", "lidm_vq_encode", "lidm_vq_quantize", "lidm_image_shape",
    "lidm_backproject", "lidm_to_uint8_image", "lidm_compact_points", "lidm_chamfer_nn", "lidm_chamfer_nn_ex",
    "lidm_chamfer_backward", "lidm_emd_forward", "lidm_emd_backward", "lidm_op_circular_conv2d", "lidm_op_conv2d_stored", "lidm_op_groupnorm", "lidm_op_qkv_attention_legacy",
    "lidm_launch_count", "lidm_profile_begin", "lidm_profile_end",
]

_lib = None


def load() -> ctypes.CDLL:
    """dlopen the in-tree library.  Raises if it has not been built (`python -m lidar_layout_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LidmError(-2, f"{LIB_PATH} not found: build it with `python -m lidar_layout_b200.build` "
                            "(no CPU fallback exists)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.lidm_last_error.restype = c_char_p
    lib.lidm_last_error.argtypes = [c_void_p]
    lib.lidm_create.argtypes = [POINTER(CConfig), POINTER(c_void_p)]
    lib.lidm_destroy.argtypes = [c_void_p]
    lib.lidm_destroy.restype = None
    lib.lidm_load_weight.argtypes = [c_void_p, c_char_p, c_void_p, c_int32, POINTER(c_int64)]
    lib.lidm_finalize_weights.argtypes = [c_void_p, c_int32]
    lib.lidm_unet_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    lib.lidm_unet_forward_cond.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32,
                                           c_void_p]
    lib.lidm_ddim_sample_cond.argtypes = [c_void_p, c_void_p, POINTER(c_int64), POINTER(c_float), c_int32, c_void_p,
                                          c_float, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                          c_float, c_void_p]
    lib.lidm_cfg_combine.argtypes = [c_void_p, c_float, c_void_p, c_int64, c_void_p]
    lib.lidm_layout_encode.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                       POINTER(c_int32), POINTER(c_void_p), c_void_p]
    lib.lidm_layout_set_cond.argtypes = [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                         POINTER(c_int32), POINTER(c_void_p), POINTER(c_int32), c_void_p]
    lib.lidm_ddim_step.argtypes = [c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_float,
                                   c_void_p, c_void_p, c_int64, c_void_p]
    lib.lidm_ddim_sample.argtypes = [c_void_p, c_void_p, POINTER(c_int64), POINTER(c_float), c_int32, c_void_p,
                                     c_float, c_void_p, c_int32, c_void_p]
    lib.lidm_vq_decode.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p]
    lib.lidm_vq_quantize.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    lib.lidm_vq_encode.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    lib.lidm_image_shape.argtypes = [c_void_p, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]
    lib.lidm_backproject.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_float, c_float, c_float, c_float, c_float,
                                     c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    lib.lidm_to_uint8_image.argtypes = [c_void_p, c_void_p, c_int64, c_void_p]
    lib.lidm_compact_points.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    lib.lidm_ddpm_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, ctypes.c_int64, c_int32, c_void_p, c_void_p, c_void_p]
    lib.lidm_chamfer_nn.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p]
    lib.lidm_chamfer_nn_ex.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_int32, c_void_p]
    lib.lidm_chamfer_backward.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p]
    lib.lidm_emd_forward.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_float, c_int32, c_void_p, c_void_p, c_void_p]
    lib.lidm_emd_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]
    lib.lidm_op_circular_conv2d.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32,
                                            c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                            c_void_p, c_void_p]
    lib.lidm_op_conv2d_stored.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32,
                                          c_int32, c_int32, c_int32, c_int32, c_void_p, c_float, c_int32, c_void_p, c_void_p,
                                          c_void_p]
    lib.lidm_op_groupnorm.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_float,
                                      c_int32, c_int32, c_void_p, c_void_p]
    lib.lidm_op_qkv_attention_legacy.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    lib.lidm_launch_count.restype = c_int64
    lib.lidm_launch_count.argtypes = []
    lib.lidm_profile_begin.argtypes = []
    lib.lidm_profile_end.argtypes = [POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(ctypes.c_double),
                                     POINTER(c_int64)]
    _lib = lib
    return lib


def check(code: int, handle=None):
    if code != 0:
        msg = load().lidm_last_error(handle)
        raise LidmError(code, (msg or b"").decode(errors="replace"))


def launch_count() -> int:
    return int(load().lidm_launch_count())


def profile_begin():
    check(load().lidm_profile_begin())


def profile_end():
    """-> dict category -> dict(ms, flops, bytes, launches)."""
    arr = lambda t: (t * 4)()
    ms, fl, by, ln = arr(ctypes.c_double), arr(ctypes.c_double), arr(ctypes.c_double), arr(c_int64)
    check(load().lidm_profile_end(ms, fl, by, ln))
    names = ["conv_gemm", "groupnorm", "attention", "other"]
    return {n: dict(ms=ms[i], flops=fl[i], bytes=by[i], launches=int(ln[i])) for i, n in enumerate(names)}
