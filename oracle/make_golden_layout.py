"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/layout_encoder.npz by running the UNMODIFIED reference
`LayoutTransformerEncoder` (lidm/modules/encoders/layout_encoder.py) from /root/reference on CPU with seeded random
weights and a synthetic layout.  Runs in the build container only (the GPU box has no /root/reference).
    python -m oracle.make_golden_layout"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/lidm/modules/encoders/layout_encoder.py"
REF_ROOT = "/root/reference"

CASES = {
    # the structure of the nuScenes layout2lidar configuration (models/lidm/nuscenes/layout2lidar/config.yaml:64-82: 13
    # layout tokens, 9 classes, attention resolutions 4/2/1 on the 8x128 feature map), narrowed from hidden 256 / output
    # 1024 / 6 layers / 8 heads to keep the fixture small
    "cfg": dict(layout_length=13, hidden_dim=64, output_dim=256, num_layers=2, num_heads=4, use_final_ln=True,
                num_classes_for_layout_object=9, mask_size_for_layout_object=32,
                used_condition_types=["obj_class", "obj_bbox", "is_valid_obj"], feature_map_size=[8, 128],
                use_positional_embedding=False, resolution_to_attention=[4, 2, 1], use_key_padding_mask=False,
                not_use_layout_fusion_module=False),
    # exercises the positional embedding and a single attention resolution.  (use_key_padding_mask=True cannot be pinned:
    # the reference builds the mask from the un-squeezed class column, shape (B, L, 1), and its attention then fails in
    # einsum on the resulting 5-D weights - layout_encoder.py:224,268,77-84; the shipped config keeps it False.)
    "posemb": dict(layout_length=6, hidden_dim=64, output_dim=96, num_layers=1, num_heads=4, use_final_ln=True,
                   num_classes_for_layout_object=5, mask_size_for_layout_object=8,
                   used_condition_types=["obj_class", "obj_bbox", "is_valid_obj"], feature_map_size=[8, 128],
                   use_positional_embedding=True, resolution_to_attention=[2], use_key_padding_mask=False,
                   not_use_layout_fusion_module=False),
}


def load_reference_module():
    spec = importlib.util.spec_from_file_location("ref_layout_encoder", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    from oracle.layout_ref import synthetic_layout
    mod = load_reference_module()
    out = {}
    # the reference constructor moves its constant patch boxes to the GPU; on this CPU-only box .cuda() is made a no-op
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for name, kw in CASES.items():
            torch.manual_seed(0)
            enc = mod.LayoutTransformerEncoder(**kw).eval()
            with torch.no_grad():
                if kw["use_positional_embedding"]:
                    enc.positional_embedding.normal_(0, 0.02)
                layout = synthetic_layout(3, kw["layout_length"], kw["num_classes_for_layout_object"], seed=1)
                res = enc(layout)
            for k, v in enc.state_dict().items():
                out[f"{name}/sd/{k}"] = v.numpy()
            out[f"{name}/layout"] = layout.numpy()
            for k, v in res.items():
                out[f"{name}/out/{k}"] = v.numpy()
        # ObjectAwareCrossAttention of the layout U-Net (lidm/modules/unets/object_cross_unet.py:380-565) on the "cfg"
        # encoder outputs: 128 channels = 2 heads of 64 (+64 positional), 2x32 feature map (attention resolution 2),
        # both norm orders; every parameter randomised (proj_out is zero-initialised in the reference)
        sys.path.insert(0, REF_ROOT)
        from lidm.modules.unets.object_cross_unet import ObjectAwareCrossAttention
        cond = {k[len("cfg/out/"):]: torch.from_numpy(v) for k, v in out.items() if k.startswith("cfg/out/")}
        for tag, norm_first, norm_obj in (("oaca", False, False), ("oaca_nf", True, True)):
            torch.manual_seed(2)
            blk = ObjectAwareCrossAttention(128, num_head_channels=64, encoder_channels=64, ds=4, resolution=[2, 32],
                                            type="input", norm_first=norm_first, norm_for_obj_embedding=norm_obj).eval()
            with torch.no_grad():
                for prm in blk.parameters():
                    prm.normal_(0, 0.08)
                x = torch.randn(3, 128, 2, 32)
                y, _ = blk(x, cond)
            for k, v in blk.state_dict().items():
                out[f"{tag}/sd/{k}"] = v.numpy()
            out[f"{tag}/x"] = x.numpy()
            out[f"{tag}/y"] = y.numpy()
    finally:
        torch.Tensor.cuda = orig_cuda
    path = os.path.join(ROOT, "tests", "golden", "layout_encoder.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB", len(out), "arrays")


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
