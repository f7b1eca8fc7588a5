"""CPU: the oracle of the layout-conditioned LiDM (oracle/layout_ref.py, SURVEY §8 f2(B) groundwork) against outputs of
the unmodified reference modules (tests/golden/layout_encoder.npz, written by oracle/make_golden_layout.py; weights are
regenerated from the fixture's shape tables by `seeded_state_dict`): bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import layout_ref as LR
from oracle.make_golden_layout import CASES, OACA, STD, UNET

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layout_encoder.npz")


@pytest.fixture(scope="module")
def gold():
    g = np.load(GOLD)
    shapes = json.loads(bytes(g["shapes_json"]).decode())
    return g, shapes


def _cond(g):
    return {k[len("cfg/out/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("cfg/out/")}


@pytest.mark.parametrize("case", sorted(CASES))
def test_layout_encoder_bit_exact(gold, case):
    g, shapes = gold
    kw = CASES[case]
    sd = LR.seeded_state_dict(shapes[case], 11, STD["enc"])
    out = LR.layout_encoder_forward(
        sd, torch.from_numpy(g[case + "/layout"]), num_layers=kw["num_layers"], num_heads=kw["num_heads"],
        used_condition_types=kw["used_condition_types"], feature_map_size=kw["feature_map_size"],
        resolution_to_attention=kw["resolution_to_attention"], use_positional_embedding=kw["use_positional_embedding"],
        use_final_ln=kw["use_final_ln"], use_key_padding_mask=kw["use_key_padding_mask"],
        not_use_layout_fusion_module=kw["not_use_layout_fusion_module"])
    want = {k[len(case) + 5:]: g[k] for k in g.files if k.startswith(case + "/out/")}
    assert set(out) == set(want)
    for k, v in want.items():
        got = out[k].numpy()
        assert got.shape == v.shape and got.dtype == v.dtype, k
        np.testing.assert_array_equal(got, v, err_msg=k)


def test_shapes_of_the_shipped_configuration(gold):
    g, _ = gold
    # 13 layout tokens, hidden 64 -> xf_out (B, 64, 13), xf_proj (B, 128); patch boxes for the 4x64, 2x32, 1x16 maps
    assert g["cfg/out/xf_out"].shape == (3, 64, 13) and g["cfg/out/xf_proj"].shape == (3, 128)
    for r, n in ((4, 4 * 64), (2, 2 * 32), (1, 16)):
        assert g[f"cfg/out/image_patch_bbox_embedding_for_resolution{r}"].shape == (3, 64, n)
        assert LR.patch_boxes([8, 128], r).shape == (n, 4)
    assert g["cfg/out/key_padding_mask"].dtype == np.bool_


@pytest.mark.parametrize("tag", sorted(OACA))
def test_object_aware_cross_attention_bit_exact(gold, tag):
    """ObjectAwareCrossAttention (object_cross_unet.py:380-565): image tokens attend to image + layout tokens with
    [content | positional] query / key halves; pinned on the reference module's output for both norm orders."""
    g, shapes = gold
    norm_first, norm_obj = OACA[tag]
    sd = LR.seeded_state_dict(shapes[tag], 12, STD["oaca"])
    y = LR.object_aware_cross_attention(sd, torch.from_numpy(g[tag + "/x"]), _cond(g), num_heads=2, resolution_rows=2,
                                        norm_first=norm_first, norm_for_obj_embedding=norm_obj)
    want = g[tag + "/y"]
    assert y.shape == want.shape == (3, 128, 2, 32)
    np.testing.assert_array_equal(y.numpy(), want)
    assert float(np.abs(want - g[tag + "/x"]).max()) > 1e-2      # the block really did something (proj_out is not zero)


def test_layout_unet_bit_exact(gold):
    """LayoutDiffusionUNetModel.forward (object_cross_unet.py:923-951): FiLM ResBlocks, ResBlock up/down-sampling,
    zero-padded convs, ObjectAwareCrossAttention in both paths and the middle; per-sample timesteps."""
    g, shapes = gold
    sd = LR.seeded_state_dict(shapes["unet"], 13, STD["unet"])
    y = LR.layout_unet_forward(
        sd, torch.from_numpy(g["unet/x"]), torch.from_numpy(g["unet/t"]), _cond(g), model_channels=UNET["model_channels"],
        channel_mult=UNET["channel_mult"], num_res_blocks=UNET["num_res_blocks"], attention_ds=UNET["attention_ds"],
        image_size=UNET["image_size"], num_head_channels=UNET["num_head_channels"],
        num_attention_blocks=UNET["num_attention_blocks"], use_scale_shift_norm=UNET["use_scale_shift_norm"])
    want = g["unet/y"]
    assert y.shape == want.shape == (3, 8, 8, 128) and np.isfinite(want).all() and float(np.abs(want).max()) > 1e-3
    np.testing.assert_array_equal(y.numpy(), want)


@pytest.mark.parametrize("name", ["layout_unet_small", "layout_unet_full"])
def test_runnable_layout_unet_fixtures(name):
    """The fixtures the CUDA path is tested against (tiny_layout / the shipped layout2lidar structure, product-seeded
    weights loaded strictly into the reference module) are reproduced by the oracle."""
    from lidar_layout_b200 import config as C
    from lidar_layout_b200.weights import UNET_PREFIX, random_state_dict
    g = np.load(os.path.join(os.path.dirname(GOLD), name + ".npz"))
    cfg = C.tiny_layout() if name.endswith("small") else C.nuscenes_layout2lidar()
    u = cfg.unet
    sd = {k[len(UNET_PREFIX):]: v for k, v in random_state_dict(cfg, 0).items() if k.startswith(UNET_PREFIX)}
    B = g["x"].shape[0]
    cond = {k[5:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("cond/")}
    cond = {k: (v.expand(B, *v.shape[1:]) if k.startswith("image_patch") else v) for k, v in cond.items()}
    y = LR.layout_unet_forward(sd, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), cond, model_channels=u.model_channels,
                               channel_mult=u.channel_mult, num_res_blocks=u.num_res_blocks,
                               attention_ds=u.attention_resolutions, image_size=u.image_size,
                               num_head_channels=u.num_head_channels, num_attention_blocks=u.num_attention_blocks)
    err = float((y - torch.from_numpy(g["eps"])).norm() / torch.from_numpy(g["eps"]).norm())
    assert err < 1e-5, err


def test_encoder_spec_and_oracle_reproduce_the_unet_fixture_conditioning():
    """layout_unet_small.npz stores the layout and the reference encoder's outputs: the product's parameter spec
    (names, shapes) + the oracle reproduce them bit for bit."""
    g = np.load(os.path.join(os.path.dirname(GOLD), "layout_unet_small.npz"))
    from oracle.make_golden_layout import enc_small_weights
    le, sd = enc_small_weights()
    out = LR.layout_encoder_forward(sd, torch.from_numpy(g["layout"]), num_layers=le.num_layers, num_heads=le.num_heads,
                                    used_condition_types=le.used_condition_types, feature_map_size=le.feature_map_size,
                                    resolution_to_attention=le.resolution_to_attention)
    for k in ("xf_proj", "xf_out", "obj_class_embedding", "obj_bbox_embedding"):
        np.testing.assert_array_equal(out[k].numpy(), g["cond/" + k], err_msg=k)
    for r in le.resolution_to_attention:
        k = f"image_patch_bbox_embedding_for_resolution{r}"
        np.testing.assert_array_equal(out[k][:1].numpy(), g["cond/" + k], err_msg=k)
