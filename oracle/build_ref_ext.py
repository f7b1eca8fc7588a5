"""TEST INFRASTRUCTURE ONLY.  Builds the reference's own CUDA extensions of the evaluation toolbox - chamfer3D
(lidm/eval/modules/chamfer3D/{chamfer_cuda.cpp,chamfer3D.cu}) and emd (lidm/eval/modules/emd/{emd.cpp,emd_cuda.cu}) - from
the sources where they lie under /root/reference into oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot),
for sm_100.  They are the checker the product's own kernels (lidar_layout_b200/csrc/eval_kernels.cu) are pinned against on
the GPU: tests/test_gpu_eval_ref.py loads them with torch.ops-free `importlib` and compares.  No reference source is copied.
    python -m oracle.build_ref_ext"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/lidm/eval/modules"
OUT = os.path.join(ROOT, "oracle", "_ref")
EXTS = {
    "ref_chamfer_3D": [f"{REF}/chamfer3D/chamfer_cuda.cpp", f"{REF}/chamfer3D/chamfer3D.cu"],
    "ref_emd": [f"{REF}/emd/emd.cpp", f"{REF}/emd/emd_cuda.cu"],
}


def build(verbose=False):
    if not os.path.isdir(REF):
        return {}
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    built = {}
    for name, srcs in EXTS.items():
        bdir = os.path.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        try:
            load(name=name, sources=srcs, build_directory=bdir, verbose=verbose, is_python_module=False,
                 extra_cuda_cflags=["-O2"], with_cuda=True)
            built[name] = os.path.join(bdir, name + ".so")
        except Exception as e:                       # an unbuildable reference extension is recorded, not fatal
            built[name] = f"FAILED: {type(e).__name__}: {str(e)[-400:]}"
    return built


def load_ext(name):
    """Import a previously built extension (on the GPU box: no /root/reference, no compiler run)."""
    import importlib.util
    import torch  # noqa: F401  (the extension links against libtorch)
    path = os.path.join(OUT, name, name + ".so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    for k, v in build(verbose="-v" in sys.argv).items():
        print(k, "->", v)
