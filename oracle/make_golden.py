"""Generate tests/golden/*.npz by executing the UNMODIFIED reference modules (this container only).

    python -m oracle.make_golden            # writes tests/golden/{kitti_uncond,tiny}.npz
    python -m oracle.make_golden --batch    # full-size fixture at B = 4 (kitti_uncond_b4.npz); also --cond / --ae / --ddpm

The synthetic state-dict (lidar_layout_b200.weights.random_state_dict, numpy PCG64 => platform
independent) is loaded with strict key/shape checking into the reference's LatentDiffusion built from
the reference's own YAML; inputs come from the same PCG64 family; outputs are what the reference
computes on CPU in fp32.  The fixtures are small (latents, one decoded image, schedule tables).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lidar_layout_b200 import config as cfgmod            # noqa: E402
from lidar_layout_b200.weights import param_spec, random_encoder_state_dict, random_state_dict  # noqa: E402
from oracle import ref_shim                               # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 0
INPUT_SEED = 1000


def inputs_for(cfg, B, n_noise, seed=INPUT_SEED):
    rng = np.random.Generator(np.random.PCG64(seed))
    C, H, W = cfg.latent_shape
    x_T = rng.standard_normal((B, C, H, W), dtype=np.float32)
    noise = rng.standard_normal((n_noise, B, C, H, W), dtype=np.float32)
    z = rng.standard_normal((B, C, H, W), dtype=np.float32)
    return x_T, noise, z


def sd_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].numpy()).tobytes()[:4096])
    return h.hexdigest()


def build_reference(cfg, yaml_dict):
    """Instantiate reference LatentDiffusion from a reference-format config dict."""
    ref_shim.install()
    from lidm.utils.misc_utils import instantiate_from_config
    y = ref_shim.to_attrdict(yaml_dict)
    y.model.params.first_stage_config.params.pop("ckpt_path", None)
    y.model.params["use_ema"] = False
    model = instantiate_from_config(y.model)
    model.eval()
    return model


def load_synthetic(model, cfg, seed=WEIGHT_SEED):
    sd = random_state_dict(cfg, seed)
    ref_sd = model.state_dict()
    # every synthetic key must exist with the same shape in the reference
    for k, v in sd.items():
        assert k in ref_sd, f"synthetic key {k} not in reference state_dict"
        assert tuple(ref_sd[k].shape) == tuple(v.shape), (k, ref_sd[k].shape, v.shape)
    # every reference tensor of the U-Net / decode side must be covered
    covered = set(sd)
    for k in ref_sd:
        if k.startswith("model.diffusion_model.") or k.startswith("first_stage_model.decoder.") \
                or k.startswith("first_stage_model.post_quant_conv.") or k.startswith("first_stage_model.quantize."):
            assert k in covered, f"reference key {k} missing from synthetic spec"
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    return sd


def tiny_yaml(cfg):
    """Reference-format config dict for a LidmConfig.  Conditioned configs name '__is_first_stage__' as the cond stage
    (the condition encoders are outside the sampling path: callers hand apply_model / DDIMSampler.sample the
    already-encoded conditioning tensor), with the conditioning key spelled out."""
    u, a = cfg.unet, cfg.ae
    y = _uncond_yaml(cfg)
    if cfg.conditioning_key is not None:
        mp = y["model"]["params"]
        mp["cond_stage_config"] = "__is_first_stage__"
        mp["conditioning_key"] = cfg.conditioning_key
        mp["cond_stage_trainable"] = False
        if u.use_spatial_transformer:
            mp["unet_config"]["params"].update(use_spatial_transformer=True, context_dim=u.context_dim,
                                               transformer_depth=u.transformer_depth)
    return y


def _uncond_yaml(cfg):
    u, a = cfg.unet, cfg.ae
    return {"model": {"target": "lidm.models.diffusion.ddpm.LatentDiffusion", "params": {
        "linear_start": cfg.linear_start, "linear_end": cfg.linear_end, "num_timesteps_cond": 1,
        "log_every_t": 200, "timesteps": cfg.timesteps, "image_size": list(cfg.image_size),
        "channels": cfg.channels, "first_stage_key": "image",
        "unet_config": {"target": "lidm.modules.diffusion.openaimodel.UNetModel", "params": {
            "image_size": list(u.image_size), "in_channels": u.in_channels, "out_channels": u.out_channels,
            "model_channels": u.model_channels, "attention_resolutions": list(u.attention_resolutions),
            "num_res_blocks": u.num_res_blocks, "channel_mult": list(u.channel_mult),
            "num_head_channels": u.num_head_channels, "lib_name": "lidm"}},
        "first_stage_config": {"target": "lidm.models.autoencoder.VQModelInterface", "params": {
            "embed_dim": a.embed_dim, "n_embed": a.n_embed, "lib_name": "lidm", "use_mask": a.use_mask,
            "ddconfig": {"double_z": False, "z_channels": a.z_channels, "in_channels": a.in_channels,
                         "out_ch": a.out_ch, "ch": a.ch, "ch_mult": list(a.ch_mult),
                         "strides": [list(s) for s in a.strides], "num_res_blocks": a.num_res_blocks,
                         "attn_levels": [], "dropout": 0.0},
            "lossconfig": {"target": "torch.nn.Identity"}}},
        "cond_stage_config": "__is_unconditional__"}}}


@torch.no_grad()
def run_reference(model, cfg, B, S_short, out):
    """Execute the reference API on the synthetic inputs and collect outputs into `out`."""
    from lidm.models.diffusion import ddim as ref_ddim
    from lidm.utils import lidar_utils as ref_lu
    x_T, noise, z = inputs_for(cfg, B, S_short + 2)
    x_T_t, z_t = torch.from_numpy(x_T), torch.from_numpy(z)
    shape = cfg.latent_shape

    # (1) one U-Net evaluation through the reference's own hook (apply_model), two timesteps
    for t_val in (501, 21):
        t = torch.full((B,), t_val, dtype=torch.long)
        out[f"eps_t{t_val}"] = model.apply_model(x_T_t, t, None).numpy()

    # (2) DDIM schedules as the sampler itself builds them
    sampler = ref_ddim.DDIMSampler(model)
    for S, eta in ((50, 0.0), (50, 1.0), (S_short, 0.0), (S_short, 1.0)):
        sampler.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=False)
        n = len(sampler.ddim_timesteps)
        tab = np.zeros((n, 4), dtype=np.float32)
        for i in range(n):
            tab[i, 0] = torch.full((1,), sampler.ddim_alphas[i]).item()
            tab[i, 1] = torch.full((1,), sampler.ddim_alphas_prev[i]).item()
            tab[i, 2] = torch.full((1,), sampler.ddim_sigmas[i]).item()
            tab[i, 3] = torch.full((1,), sampler.ddim_sqrt_one_minus_alphas[i]).item()
        out[f"ddim_S{S}_eta{int(eta)}_timesteps"] = np.asarray(sampler.ddim_timesteps, dtype=np.int64)
        out[f"ddim_S{S}_eta{int(eta)}_table"] = tab

    # (3) short free-running DDIM, eta=0, with a spy on apply_model (teacher-forcing data)
    rec = []
    orig_apply = model.apply_model

    def spy(x, t, c, *a, **k):
        e = orig_apply(x, t, c, *a, **k)
        rec.append((x.clone(), t.clone(), e.clone()))
        return e

    model.apply_model = spy
    samples, inter = sampler.sample(S_short, batch_size=B, shape=shape, eta=0.0, x_T=x_T_t.clone(), verbose=False)
    out["ddim_eta0_final"] = samples.numpy()
    out["ddim_eta0_xt"] = np.stack([r[0].numpy() for r in rec])
    out["ddim_eta0_t"] = np.stack([r[1].numpy() for r in rec])
    out["ddim_eta0_eps"] = np.stack([r[2].numpy() for r in rec])

    # (4) eta=1 with injected noise (ddim.py:202 draws noise_like once per step)
    rec.clear()
    it = iter(list(noise))
    orig_noise_like = ref_ddim.noise_like
    ref_ddim.noise_like = lambda shape_, device, repeat=False: torch.from_numpy(next(it))
    try:
        samples1, _ = sampler.sample(S_short, batch_size=B, shape=shape, eta=1.0, x_T=x_T_t.clone(), verbose=False)
    finally:
        ref_ddim.noise_like = orig_noise_like
        model.apply_model = orig_apply
    out["ddim_eta1_final"] = samples1.numpy()

    # (5) first stage decode, quantised and not; VQ indices through the quantiser itself
    out["decode_q"] = model.decode_first_stage(z_t).numpy()
    out["decode_nq"] = model.decode_first_stage(z_t, force_not_quantize=True).numpy()
    _, _, (_, _, idx) = model.first_stage_model.quantize(z_t)
    out["vq_idx"] = idx.numpy().astype(np.int32)

    # (6) back-projection of the decoded image exactly as scripts/sample.py:29-35 does
    ds = dict(fov=list(cfg.dataset.fov), depth_range=list(cfg.dataset.depth_range),
              depth_scale=cfg.dataset.depth_scale, log_scale=cfg.dataset.log_scale)
    # synthetic range image with a realistic spread (decoded random-init images sit near 0)
    rng = np.random.Generator(np.random.PCG64(INPUT_SEED + 1))
    H, W = out["decode_q"].shape[-2:]
    img = np.clip(rng.standard_normal((H, W), dtype=np.float32) * 0.6, -1.2, 1.2).astype(np.float32)
    out["bp_img"] = img
    unit = (np.clip(img, -1.0, 1.0) + 1.0) / 2.0
    pcd, _, _ = ref_lu.range2pcd(unit, **ds)
    out["bp_pcd"] = pcd
    out["bp_xyz"] = ref_lu.range2xyz(unit, **ds)


def cond_inputs_for(cfg, B, L, seed=INPUT_SEED + 7):
    """Synthetic conditioning: crossattn -> (context, unconditional context) (B, L, context_dim);
    concat -> (c_concat, None) (B, in_channels - channels, H, W)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    C, H, W = cfg.latent_shape
    if cfg.conditioning_key == "crossattn":
        d = cfg.unet.context_dim
        return rng.standard_normal((B, L, d), dtype=np.float32), rng.standard_normal((B, L, d), dtype=np.float32)
    if cfg.conditioning_key == "concat":
        return rng.standard_normal((B, cfg.unet.in_channels - C, H, W), dtype=np.float32), None
    raise ValueError(cfg.conditioning_key)


CFG_SCALE = 2.5


@torch.no_grad()
def run_reference_cond(model, cfg, B, S_short, L, out, full=True):
    """Conditioned sampling through the reference API: apply_model with a conditioning tensor, DDIMSampler.sample
    with `conditioning`, and (crossattn) classifier-free guidance (ddim.py:173-180)."""
    from lidm.models.diffusion import ddim as ref_ddim
    x_T, noise, _ = inputs_for(cfg, B, S_short + 2)
    cond, uncond = cond_inputs_for(cfg, B, L)
    x_T_t, c_t = torch.from_numpy(x_T), torch.from_numpy(cond)
    uc_t = torch.from_numpy(uncond) if uncond is not None else None
    out["L"] = np.int64(L)
    for t_val in (501, 21):
        t = torch.full((B,), t_val, dtype=torch.long)
        out[f"eps_t{t_val}"] = model.apply_model(x_T_t, t, c_t).numpy()
    if not full:
        return
    sampler = ref_ddim.DDIMSampler(model)
    rec = []
    orig_apply = model.apply_model

    def spy(x, t, c, *a, **k):
        e = orig_apply(x, t, c, *a, **k)
        rec.append((x.clone(), t.clone(), e.clone()))
        return e

    model.apply_model = spy
    try:
        samples, _ = sampler.sample(S_short, batch_size=B, shape=cfg.latent_shape, conditioning=c_t, eta=0.0,
                                    x_T=x_T_t.clone(), verbose=False)
        out["ddim_eta0_final"] = samples.numpy()
        out["ddim_eta0_xt"] = np.stack([r[0].numpy() for r in rec])
        out["ddim_eta0_eps"] = np.stack([r[2].numpy() for r in rec])
        if uc_t is not None:
            rec.clear()
            samples, _ = sampler.sample(S_short, batch_size=B, shape=cfg.latent_shape, conditioning=c_t, eta=0.0,
                                        x_T=x_T_t.clone(), verbose=False, unconditional_guidance_scale=CFG_SCALE,
                                        unconditional_conditioning=uc_t)
            out["ddim_cfg_final"] = samples.numpy()
            out["cfg_scale"] = np.float32(CFG_SCALE)
            # the (2B) batched [uncond | cond] eps of the first step, as apply_model returned it
            out["ddim_cfg_eps2b_step0"] = rec[0][2].numpy()
            # eta=1 + guidance + injected noise
            it = iter(list(noise))
            orig_noise_like = ref_ddim.noise_like
            ref_ddim.noise_like = lambda shape_, device, repeat=False: torch.from_numpy(next(it))
            try:
                samples, _ = sampler.sample(S_short, batch_size=B, shape=cfg.latent_shape, conditioning=c_t, eta=1.0,
                                            x_T=x_T_t.clone(), verbose=False,
                                            unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=uc_t)
            finally:
                ref_ddim.noise_like = orig_noise_like
            out["ddim_cfg_eta1_final"] = samples.numpy()
    finally:
        model.apply_model = orig_apply


def make_cond(name, cfg, yaml_dict, B, S_short, L, full=True):
    model = build_reference(cfg, yaml_dict)
    sd = load_synthetic(model, cfg)
    out = {"weights_digest": np.frombuffer(sd_digest(sd).encode(), dtype=np.uint8),
           "weight_seed": np.int64(WEIGHT_SEED), "input_seed": np.int64(INPUT_SEED),
           "B": np.int64(B), "S_short": np.int64(S_short)}
    run_reference_cond(model, cfg, B, S_short, L, out, full=full)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; keys={sorted(out)}")


def make(name, cfg, yaml_dict, B, S_short):
    model = build_reference(cfg, yaml_dict)
    sd = load_synthetic(model, cfg)
    out = {"weights_digest": np.frombuffer(sd_digest(sd).encode(), dtype=np.uint8),
           "weight_seed": np.int64(WEIGHT_SEED), "input_seed": np.int64(INPUT_SEED),
           "B": np.int64(B), "S_short": np.int64(S_short)}
    run_reference(model, cfg, B, S_short, out)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; keys={sorted(out)}")


def ae_images_for(cfg, B, seed=INPUT_SEED + 3):
    """Synthetic range images in the dataset's value range: log-depth-like values in [-1, 1], 20 % empty pixels (-1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    H, W = cfg.dataset.size
    x = np.clip(rng.standard_normal((B, cfg.ae.in_channels, H, W), dtype=np.float32) * 0.5, -1.0, 1.0)
    x[rng.random((B, cfg.ae.in_channels, H, W)) < 0.2] = -1.0
    return x.astype(np.float32)


@torch.no_grad()
def make_ae(name, cfg, yaml_dict, B):
    """First-stage autoencoder: VQModelInterface.encode and the encode -> decode round trip
    through the reference's own LatentDiffusion.encode_first_stage / decode_first_stage."""
    model = build_reference(cfg, yaml_dict)
    sd = load_synthetic(model, cfg)
    esd = random_encoder_state_dict(cfg, WEIGHT_SEED)
    ref_sd = model.state_dict()
    for k, v in esd.items():
        assert k in ref_sd and tuple(ref_sd[k].shape) == tuple(v.shape), k
    for k in ref_sd:
        if k.startswith("first_stage_model.encoder.") or k.startswith("first_stage_model.quant_conv."):
            assert k in esd, f"reference key {k} missing from the encoder spec"
    model.load_state_dict(esd, strict=False)
    x = torch.from_numpy(ae_images_for(cfg, B))
    out = {"B": np.int64(B), "weight_seed": np.int64(WEIGHT_SEED),
           "weights_digest": np.frombuffer(sd_digest({**sd, **esd}).encode(), dtype=np.uint8)}
    z = model.encode_first_stage(x)
    out["encode"] = z.numpy()
    out["encoding_scaled"] = model.get_first_stage_encoding(z).numpy()
    out["recon_q"] = model.decode_first_stage(z).numpy()
    out["recon_nq"] = model.decode_first_stage(z, force_not_quantize=True).numpy()
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; keys={sorted(out)}")


def main_ae():
    torch.set_num_threads(os.cpu_count())
    import yaml
    tiny = cfgmod.tiny()
    make_ae("tiny_ae", tiny, tiny_yaml(tiny), B=2)
    with open(os.path.join(ref_shim.REFERENCE_ROOT, "models/lidm/kitti/uncond/config.yaml")) as f:
        y = yaml.safe_load(f)
    make_ae("kitti_ae", cfgmod.from_reference_dict(y), y, B=1)


@torch.no_grad()
def make_ddpm(name, cfg, yaml_dict, B, steps):
    """Ancestral sampling: LatentDiffusion.sample(..., timesteps=steps) with the per-step
    noise of p_sample (ddpm.py:1106) injected, recorded with a spy on apply_model."""
    model = build_reference(cfg, yaml_dict)
    from lidm.models.diffusion import ddpm as ref_ddpm
    sd = load_synthetic(model, cfg)
    x_T, noise, _ = inputs_for(cfg, B, steps)
    rec = []
    orig_apply = model.apply_model

    def spy(x, t, c, *a, **k):
        e = orig_apply(x, t, c, *a, **k)
        rec.append((x.clone(), t.clone(), e.clone()))
        return e

    model.apply_model = spy
    it = iter(list(noise))
    orig_noise_like = ref_ddpm.noise_like
    ref_ddpm.noise_like = lambda shape_, device, repeat=False: torch.from_numpy(next(it))
    try:
        out_img = model.sample(None, batch_size=B, x_T=torch.from_numpy(x_T).clone(), timesteps=steps)
    finally:
        ref_ddpm.noise_like = orig_noise_like
        model.apply_model = orig_apply
    out = {"B": np.int64(B), "steps": np.int64(steps), "weight_seed": np.int64(WEIGHT_SEED),
           "weights_digest": np.frombuffer(sd_digest(sd).encode(), dtype=np.uint8),
           "final": out_img.numpy(), "xt": np.stack([r[0].numpy() for r in rec]),
           "t": np.stack([r[1].numpy() for r in rec]), "eps": np.stack([r[2].numpy() for r in rec])}
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; keys={sorted(out)}")


@torch.no_grad()
def make_inpaint(name, cfg, yaml_dict, B, S):
    """DDIM inpainting blend (ddim.py:146-149): img = q_sample(x0, t) * mask + (1 - mask) * img before every step;
    q_sample's randn_like draws are injected."""
    model = build_reference(cfg, yaml_dict)
    from lidm.models.diffusion import ddim as ref_ddim
    sd = load_synthetic(model, cfg)
    x_T, noise, x0 = inputs_for(cfg, B, S)
    rng = np.random.Generator(np.random.PCG64(INPUT_SEED + 11))
    mask = (rng.random((B, 1) + tuple(cfg.latent_shape[1:])) < 0.5).astype(np.float32)
    it = iter(list(noise))
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: torch.from_numpy(next(it))
    try:
        sampler = ref_ddim.DDIMSampler(model)
        out_img, _ = sampler.sample(S, batch_size=B, shape=cfg.latent_shape, eta=0.0, x_T=torch.from_numpy(x_T).clone(),
                                    mask=torch.from_numpy(mask), x0=torch.from_numpy(x0), verbose=False)
    finally:
        torch.randn_like = orig
    out = {"B": np.int64(B), "S": np.int64(S), "mask": mask, "final": out_img.numpy(),
           "weights_digest": np.frombuffer(sd_digest(sd).encode(), dtype=np.uint8)}
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; keys={sorted(out)}")


def main_ddpm():
    torch.set_num_threads(os.cpu_count())
    tiny = cfgmod.tiny()
    make_ddpm("tiny_ddpm", tiny, tiny_yaml(tiny), B=2, steps=5)
    make_inpaint("tiny_inpaint", tiny, tiny_yaml(tiny), B=2, S=4)


def main_cond():
    """Conditioned fixtures."""
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    for name, cfg, L in (("tiny_crossattn", cfgmod.tiny(cond="crossattn"), 5), ("tiny_concat", cfgmod.tiny(cond="concat"), 0)):
        make_cond(name, cfg, tiny_yaml(cfg), B=2, S_short=4, L=L)
    # full-size cross-attention U-Net (models/lidm/kitti/cam2lidar): eps only, two context lengths (camera: 4 tokens,
    # text: 77 tokens)
    full = cfgmod.kitti_cam2lidar()
    make_cond("kitti_cam2lidar_L4", full, tiny_yaml(full), B=1, S_short=0, L=4, full=False)
    make_cond("kitti_cam2lidar_L77", full, tiny_yaml(full), B=1, S_short=0, L=77, full=False)


@torch.no_grad()
def make_batch(name, cfg, yaml_dict, B):
    """Full-size parity at a batch > 1: apply_model with per-sample timesteps and
    decode_first_stage (quantised and not) of the unmodified reference at batch B, on inputs of their own seed."""
    model = build_reference(cfg, yaml_dict)
    sd = load_synthetic(model, cfg)
    x_T, _, z = inputs_for(cfg, B, 1, seed=INPUT_SEED + 21)
    t = np.asarray([501, 21, 900, 333, 7, 650, 120, 999][:B], dtype=np.int64)
    out = {"B": np.int64(B), "weight_seed": np.int64(WEIGHT_SEED), "input_seed": np.int64(INPUT_SEED + 21),
           "weights_digest": np.frombuffer(sd_digest(sd).encode(), dtype=np.uint8), "t": t}
    out["eps"] = model.apply_model(torch.from_numpy(x_T), torch.from_numpy(t), None).numpy()
    out["decode_q"] = model.decode_first_stage(torch.from_numpy(z)).numpy()
    out["decode_nq"] = model.decode_first_stage(torch.from_numpy(z), force_not_quantize=True).numpy()
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB; keys={sorted(out)}")


def main_batch():
    torch.set_num_threads(os.cpu_count())
    import yaml
    with open(os.path.join(ref_shim.REFERENCE_ROOT, "models/lidm/kitti/uncond/config.yaml")) as f:
        y = yaml.safe_load(f)
    make_batch("kitti_uncond_b4", cfgmod.from_reference_dict(y), y, B=4)


def main():
    if "--batch" in sys.argv:
        return main_batch()
    if "--cond" in sys.argv:
        return main_cond()
    if "--ae" in sys.argv:
        return main_ae()
    if "--ddpm" in sys.argv:
        return main_ddpm()
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    import yaml
    tiny = cfgmod.tiny()
    make("tiny", tiny, tiny_yaml(tiny), B=2, S_short=4)
    with open(os.path.join(ref_shim.REFERENCE_ROOT, "models/lidm/kitti/uncond/config.yaml")) as f:
        y = yaml.safe_load(f)
    full = cfgmod.from_reference_dict(y)
    make("kitti_uncond", full, y, B=1, S_short=4)


if __name__ == "__main__":
    main()
