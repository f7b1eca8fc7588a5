"""CPU: the layout-encoder oracle (oracle/layout_ref.py, SURVEY §8 f2(B) groundwork) against outputs of the unmodified
reference module (tests/golden/layout_encoder.npz, written by oracle/make_golden_layout.py): bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import layout_ref as LR
from oracle.make_golden_layout import CASES

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layout_encoder.npz")


@pytest.mark.parametrize("case", sorted(CASES))
def test_layout_encoder_bit_exact(case):
    g = np.load(GOLD)
    kw = CASES[case]
    sd = {k[len(case) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(case + "/sd/")}
    layout = torch.from_numpy(g[case + "/layout"])
    out = LR.layout_encoder_forward(
        sd, layout, num_layers=kw["num_layers"], num_heads=kw["num_heads"], used_condition_types=kw["used_condition_types"],
        feature_map_size=kw["feature_map_size"], resolution_to_attention=kw["resolution_to_attention"],
        use_positional_embedding=kw["use_positional_embedding"], use_final_ln=kw["use_final_ln"],
        use_key_padding_mask=kw["use_key_padding_mask"], not_use_layout_fusion_module=kw["not_use_layout_fusion_module"])
    want = {k[len(case) + 5:]: g[k] for k in g.files if k.startswith(case + "/out/")}
    assert set(out) == set(want)
    for k, v in want.items():
        got = out[k].numpy()
        assert got.shape == v.shape and got.dtype == v.dtype, k
        np.testing.assert_array_equal(got, v, err_msg=k)


def test_shapes_of_the_shipped_configuration():
    g = np.load(GOLD)
    # 13 layout tokens, hidden 64 -> xf_out (B, 64, 13), xf_proj (B, 256); patch boxes for the 4x64, 2x32, 1x16 maps
    assert g["cfg/out/xf_out"].shape == (3, 64, 13) and g["cfg/out/xf_proj"].shape == (3, 256)
    for r, n in ((4, 4 * 64), (2, 2 * 32), (1, 16)):
        assert g[f"cfg/out/image_patch_bbox_embedding_for_resolution{r}"].shape == (3, 64, n)
        assert LR.patch_boxes([8, 128], r).shape == (n, 4)
    assert g["cfg/out/key_padding_mask"].dtype == np.bool_


@pytest.mark.parametrize("tag,norm_first,norm_obj", [("oaca", False, False), ("oaca_nf", True, True)])
def test_object_aware_cross_attention_bit_exact(tag, norm_first, norm_obj):
    """ObjectAwareCrossAttention (object_cross_unet.py:380-565): image tokens attend to image + layout tokens with
    [content | positional] query / key halves; pinned on the reference module's output for both norm orders."""
    g = np.load(GOLD)
    sd = {k[len(tag) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + "/sd/")}
    cond = {k[len("cfg/out/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("cfg/out/")}
    y = LR.object_aware_cross_attention(sd, torch.from_numpy(g[tag + "/x"]), cond, num_heads=2, resolution_rows=2,
                                        norm_first=norm_first, norm_for_obj_embedding=norm_obj)
    want = g[tag + "/y"]
    assert y.shape == want.shape == (3, 128, 2, 32)
    np.testing.assert_array_equal(y.numpy(), want)
    assert float(np.abs(want - g[tag + "/x"]).max()) > 1e-2      # the block really did something (proj_out randomised)
