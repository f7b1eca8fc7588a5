"""TEST INFRASTRUCTURE ONLY.  Writes tests/golden/layout_encoder.npz by running the UNMODIFIED reference modules of the
layout-conditioned LiDM from /root/reference on CPU: `LayoutTransformerEncoder`
(lidm/modules/encoders/layout_encoder.py), `ObjectAwareCrossAttention` and `LayoutDiffusionUNetModel`
(lidm/modules/unets/object_cross_unet.py), with seeded weights (oracle.layout_ref.seeded_state_dict: the fixture stores
only the {name: shape} tables) and synthetic inputs.  Runs in the build container only (the GPU box has no
/root/reference).
    python -m oracle.make_golden_layout"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = "/root/reference"
REF = REF_ROOT + "/lidm/modules/encoders/layout_encoder.py"

CASES = {
    # the structure of the nuScenes layout2lidar configuration (models/lidm/nuscenes/layout2lidar/config.yaml:64-82: 13
    # layout tokens, 9 classes, attention resolutions 4/2/1 on the 8x128 feature map), narrowed from hidden 256 / output
    # 1024 / 6 layers / 8 heads to keep the run small
    "cfg": dict(layout_length=13, hidden_dim=64, output_dim=128, num_layers=2, num_heads=4, use_final_ln=True,
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
# ObjectAwareCrossAttention alone: 128 channels = 2 heads of 64 (+64 positional), 2x32 feature map, both norm orders
OACA = {"oaca": (False, False), "oaca_nf": (True, True)}
# the whole denoiser at a small width (the shipped one: model_channels 256, mult [1,2,4], 2 res blocks, attention_ds [8,4,2])
UNET = dict(image_size=[8, 128], use_fp16=False, use_scale_shift_norm=True, in_channels=8, out_channels=8, model_channels=32,
            encoder_channels=64, num_head_channels=64, num_heads=-1, num_heads_upsample=-1, num_res_blocks=1,
            num_attention_blocks=1, resblock_updown=True, attention_ds=[2], channel_mult=[1, 2], dropout=0.1,
            use_checkpoint=False, use_positional_embedding_for_attention=True,
            attention_block_type="ObjectAwareCrossAttention")
STD = {"enc": 0.05, "oaca": 0.08, "unet": 0.05}


def load_reference_module():
    spec = importlib.util.spec_from_file_location("ref_layout_encoder", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _seed_module(module, seed, std):
    from oracle.layout_ref import seeded_state_dict
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    module.load_state_dict(seeded_state_dict(shapes, seed, std))
    return shapes


def main():
    from oracle.layout_ref import synthetic_layout
    mod = load_reference_module()
    out, shapes = {}, {}
    # the reference constructor moves its constant patch boxes to the GPU; on this CPU-only box .cuda() is made a no-op
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        cond = None
        for name, kw in CASES.items():
            enc = mod.LayoutTransformerEncoder(**kw).eval()
            if kw["use_positional_embedding"]:
                with torch.no_grad():
                    enc.positional_embedding.zero_()          # torch.empty in the reference: give it defined contents
            shapes[name] = _seed_module(enc, 11, STD["enc"])
            with torch.no_grad():
                layout = synthetic_layout(3, kw["layout_length"], kw["num_classes_for_layout_object"], seed=1)
                res = enc(layout)
            out[f"{name}/layout"] = layout.numpy()
            for k, v in res.items():
                out[f"{name}/out/{k}"] = v.numpy()
            if name == "cfg":
                cond = res
        sys.path.insert(0, REF_ROOT)
        from lidm.modules.unets.object_cross_unet import LayoutDiffusionUNetModel, ObjectAwareCrossAttention
        for tag, (norm_first, norm_obj) in OACA.items():
            blk = ObjectAwareCrossAttention(128, num_head_channels=64, encoder_channels=64, ds=4, resolution=[2, 32],
                                            type="input", norm_first=norm_first, norm_for_obj_embedding=norm_obj).eval()
            shapes[tag] = _seed_module(blk, 12, STD["oaca"])
            with torch.no_grad():
                x = torch.randn(3, 128, 2, 32, generator=torch.Generator().manual_seed(2))
                y, _ = blk(x, cond)
            out[f"{tag}/x"], out[f"{tag}/y"] = x.numpy(), y.numpy()
        net = LayoutDiffusionUNetModel(**UNET).eval()
        shapes["unet"] = _seed_module(net, 13, STD["unet"])
        with torch.no_grad():
            x = torch.randn(3, 8, 8, 128, generator=torch.Generator().manual_seed(3))
            t = torch.tensor([3, 500, 977])
            y = net(x, t, cond)
        out["unet/x"], out["unet/t"], out["unet/y"] = x.numpy(), t.numpy(), y.numpy()
    finally:
        torch.Tensor.cuda = orig_cuda
    out["shapes_json"] = np.frombuffer(json.dumps({k: {n: list(s) for n, s in v.items()} for k, v in shapes.items()}).encode(),
                                       dtype=np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "layout_encoder.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB", len(out), "arrays")


# --------------------------------------------------------------------------------------------------------------------
# Runnable-width fixtures for the CUDA path (python -m oracle.make_golden_layout --unet): the reference
# LayoutDiffusionUNetModel built from the product's config objects (tiny_layout: the shipped structure at 64 channels;
# nuscenes_layout2lidar: the shipped models/lidm/nuscenes/layout2lidar/config.yaml), loaded STRICTLY with the product's
# seeded state-dict (lidar_layout_b200.weights.random_state_dict: pins every key and shape), conditioned on the outputs of
# the reference LayoutTransformerEncoder (seeded weights) for a synthetic layout.  Stored: x, t, the conditioning dict, eps.
ENC_FULL = dict(layout_length=13, hidden_dim=256, output_dim=1024, num_layers=6, num_heads=8, use_final_ln=True,
                num_classes_for_layout_object=9, mask_size_for_layout_object=32,
                used_condition_types=["obj_class", "obj_bbox", "is_valid_obj"], feature_map_size=[8, 128],
                use_positional_embedding=False, resolution_to_attention=[4, 2, 1], use_key_padding_mask=False,
                not_use_layout_fusion_module=False)
ENC_SMALL = dict(ENC_FULL, hidden_dim=64, output_dim=256, num_layers=2, num_heads=4)


def enc_small_weights():
    """The seeded encoder weights of the layout_unet_small fixture (seed 21, as main_unet draws them), rebuilt from the
    product's own parameter spec: (LayoutEncoderConfig, {name without the cond_stage_model. prefix: tensor})."""
    from lidar_layout_b200 import config as C
    from lidar_layout_b200.weights import COND_PREFIX, layout_encoder_param_spec
    from oracle.layout_ref import seeded_state_dict
    le = C.tiny_layout().layout_encoder
    shapes = {k[len(COND_PREFIX):]: shp for k, (shp, _) in layout_encoder_param_spec(le).items()}
    return le, seeded_state_dict(shapes, 21, STD["enc"])


def main_unet():
    from lidar_layout_b200 import config as C
    from lidar_layout_b200.weights import UNET_PREFIX, random_state_dict
    from oracle.layout_ref import synthetic_layout
    mod = load_reference_module()
    sys.path.insert(0, REF_ROOT)
    from lidm.modules.unets.object_cross_unet import LayoutDiffusionUNetModel
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.set_num_threads(os.cpu_count())
    try:
        for name, cfg, enc_kw, B in (("layout_unet_small", C.tiny_layout(), ENC_SMALL, 3),
                                     ("layout_unet_full", C.nuscenes_layout2lidar(), ENC_FULL, 2)):
            u = cfg.unet
            enc = mod.LayoutTransformerEncoder(**enc_kw).eval()
            _seed_module(enc, 21, STD["enc"])
            net = LayoutDiffusionUNetModel(
                image_size=list(u.image_size), use_fp16=False, use_scale_shift_norm=True, in_channels=u.in_channels,
                out_channels=u.out_channels, model_channels=u.model_channels, encoder_channels=u.encoder_channels,
                num_head_channels=u.num_head_channels, num_heads=-1, num_heads_upsample=-1, num_res_blocks=u.num_res_blocks,
                num_attention_blocks=u.num_attention_blocks, resblock_updown=True, attention_ds=list(u.attention_resolutions),
                channel_mult=list(u.channel_mult), dropout=0.1, use_checkpoint=False,
                use_positional_embedding_for_attention=True, attention_block_type="ObjectAwareCrossAttention").eval()
            sd = random_state_dict(cfg, 0)
            usd = {k[len(UNET_PREFIX):]: v for k, v in sd.items() if k.startswith(UNET_PREFIX)}
            net.load_state_dict(usd, strict=True)            # every key and shape of the product's spec is the reference's
            with torch.no_grad():
                layout = synthetic_layout(B, 13, 9, seed=5)
                cond = enc(layout)
                g = torch.Generator().manual_seed(6)
                x = torch.randn(B, u.in_channels, *u.image_size, generator=g)
                t = torch.tensor([3, 500, 977][:B])
                y = net(x, t, cond)
            out = {"x": x.numpy(), "t": t.numpy(), "eps": y.numpy(), "layout": layout.numpy()}
            for k, v in cond.items():
                if k.startswith("image_patch_bbox_embedding"):
                    assert torch.equal(v[:1].expand_as(v), v)
                    v = v[:1]                                   # batch-broadcast by construction: store one row
                out["cond/" + k] = v.numpy()
            path = os.path.join(ROOT, "tests", "golden", name + ".npz")
            np.savez_compressed(path, **out)
            print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB", "eps rms", float(y.pow(2).mean().sqrt()))
    finally:
        torch.Tensor.cuda = orig_cuda


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    if "--unet" in sys.argv:
        main_unet()
    else:
        main()
