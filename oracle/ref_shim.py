"""Import shim that lets the UNMODIFIED reference (/root/reference) run in this container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (lidar_layout_b200/) may import this
file.  It is used by oracle/make_golden.py (to generate tests/golden/*.npz from the real
reference modules) and by tests that run *here* to pin oracle/torch_ref.py against the
reference.  /root/reference does not exist on the GPU box, so every use is guarded by
`reference_available()`.

What is stubbed:
  * pytorch_lightning.LightningModule -> nn.Module with .device/.log/.log_dict
  * pytorch_lightning.utilities.distributed.rank_zero_only -> identity decorator
  * fvdb, fvdb.nn.VDBTensor -> empty classes (only used in isinstance checks)
  * taming.modules.vqvae.quantize.VectorQuantizer2 -> restatement of taming's quantiser,
    arithmetic as lidm/models/ae/vq.py:66-106 plus the b c h w -> b h w c permute that the
    1-D copy drops (taming-transformers is installed unpinned from git master by
    init/create_env.sh:15 and is not vendored in the reference tree)
  * omegaconf.listconfig.ListConfig -> list subclass
  * sys.modules['lidm.models.autoencoder'] alias (stale config target,
    models/lidm/kitti/uncond/config.yaml:27)
  * DDIMSampler.register_buffer -> plain setattr (lidm/models/diffusion/ddim.py:20-24 forces .cuda())
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LIDM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lidm"))


class AttrDict(dict):
    """dict with attribute access (stands in for OmegaConf DictConfig)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attrdict(o):
    if isinstance(o, dict):
        return AttrDict({k: to_attrdict(v) for k, v in o.items()})
    if isinstance(o, list):
        return [to_attrdict(v) for v in o]
    return o


_installed = False


def install():
    """Install the stubs and put the reference on sys.path.  Idempotent."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch
    import torch.nn as nn

    # ---- pytorch_lightning ------------------------------------------------------------
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    pl_util = types.ModuleType("pytorch_lightning.utilities")
    pl_dist = types.ModuleType("pytorch_lightning.utilities.distributed")
    pl_dist.rank_zero_only = lambda f: f
    pl_util.distributed = pl_dist
    pl.utilities = pl_util
    sys.modules.setdefault("pytorch_lightning", pl)
    sys.modules.setdefault("pytorch_lightning.utilities", pl_util)
    sys.modules.setdefault("pytorch_lightning.utilities.distributed", pl_dist)

    # ---- fvdb -------------------------------------------------------------------------
    fvdb = types.ModuleType("fvdb")
    fvnn = types.ModuleType("fvdb.nn")

    class VDBTensor:  # only isinstance() targets
        pass

    class GridBatch:
        pass

    class JaggedTensor:
        pass

    fvnn.VDBTensor = VDBTensor
    fvdb.nn = fvnn
    fvdb.GridBatch = GridBatch
    fvdb.JaggedTensor = JaggedTensor
    sys.modules.setdefault("fvdb", fvdb)
    sys.modules.setdefault("fvdb.nn", fvnn)

    # ---- omegaconf.listconfig ---------------------------------------------------------
    if "omegaconf" not in sys.modules:
        try:
            import omegaconf  # noqa: F401
        except Exception:
            oc = types.ModuleType("omegaconf")
            ocl = types.ModuleType("omegaconf.listconfig")

            class ListConfig(list):
                pass

            ocl.ListConfig = ListConfig
            oc.listconfig = ocl
            oc.ListConfig = ListConfig
            sys.modules["omegaconf"] = oc
            sys.modules["omegaconf.listconfig"] = ocl

    # ---- taming VectorQuantizer2 ------------------------------------------------------
    taming = types.ModuleType("taming")
    tm = types.ModuleType("taming.modules")
    tv = types.ModuleType("taming.modules.vqvae")
    tq = types.ModuleType("taming.modules.vqvae.quantize")

    class VectorQuantizer2(nn.Module):
        """taming-transformers VectorQuantizer2 (eval-path arithmetic).

        Follows lidm/models/ae/vq.py:16-40 (init) and :66-106 (forward) with taming's
        channel-last rearrangement restored around the flatten.
        """

        def __init__(self, n_e, e_dim, beta, remap=None, unknown_index="random",
                     sane_index_shape=False, legacy=True):
            super().__init__()
            assert remap is None
            self.n_e, self.e_dim, self.beta, self.legacy = n_e, e_dim, beta, legacy
            self.embedding = nn.Embedding(self.n_e, self.e_dim)
            self.embedding.weight.data.uniform_(-1.0 / self.n_e, 1.0 / self.n_e)
            self.re_embed = n_e
            self.sane_index_shape = sane_index_shape

        def forward(self, z, temp=None, rescale_logits=False, return_logits=False):
            z = z.permute(0, 2, 3, 1).contiguous()                      # b c h w -> b h w c
            z_flattened = z.view(-1, self.e_dim)
            d = torch.sum(z_flattened ** 2, dim=1, keepdim=True) + \
                torch.sum(self.embedding.weight ** 2, dim=1) - 2 * \
                torch.einsum('bd,dn->bn', z_flattened, self.embedding.weight.t())
            min_encoding_indices = torch.argmin(d, dim=1)
            z_q = self.embedding(min_encoding_indices).view(z.shape)
            loss = torch.mean((z_q.detach() - z) ** 2) + self.beta * torch.mean((z_q - z.detach()) ** 2)
            z_q = z + (z_q - z).detach()
            z_q = z_q.permute(0, 3, 1, 2).contiguous()                   # b h w c -> b c h w
            if self.sane_index_shape:
                min_encoding_indices = min_encoding_indices.reshape(z_q.shape[0], z_q.shape[2], z_q.shape[3])
            return z_q, loss, (None, None, min_encoding_indices)

        def get_codebook_entry(self, indices, shape):
            z_q = self.embedding(indices)
            if shape is not None:
                z_q = z_q.view(shape).permute(0, 3, 1, 2).contiguous()
            return z_q

    tq.VectorQuantizer2 = VectorQuantizer2
    taming.modules = tm
    tm.vqvae = tv
    tv.quantize = tq
    sys.modules.setdefault("taming", taming)
    sys.modules.setdefault("taming.modules", tm)
    sys.modules.setdefault("taming.modules.vqvae", tv)
    sys.modules.setdefault("taming.modules.vqvae.quantize", tq)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # stale config targets
    import lidm.models.ae.autoencoder as _ae
    sys.modules["lidm.models.autoencoder"] = _ae

    # CPU-friendly DDIM buffers
    from lidm.models.diffusion import ddim as _ddim

    def _register_buffer(self, name, attr):
        setattr(self, name, attr)

    _ddim.DDIMSampler.register_buffer = _register_buffer
    _installed = True


def load_yaml_config(rel_path):
    """yaml.safe_load + attribute dict, path relative to the reference root."""
    import yaml
    with open(os.path.join(REFERENCE_ROOT, rel_path)) as f:
        return to_attrdict(yaml.safe_load(f))


def build_reference_lidm(config_rel="models/lidm/kitti/uncond/config.yaml", use_ema=False):
    """Instantiate the reference LatentDiffusion from its own YAML (random init, CPU)."""
    install()
    from lidm.utils.misc_utils import instantiate_from_config
    cfg = load_yaml_config(config_rel)
    params = cfg.model.params
    params.first_stage_config.params.pop("ckpt_path", None)
    params.pop("ckpt_path", None)
    params["use_ema"] = use_ema
    model = instantiate_from_config(cfg.model)
    model.eval()
    return model, cfg
