"""Mirror of the reference's point-cloud distance modules (forward and backward): `emdModule` (reference
lidm/eval/modules/emd/emd_module.py:47-98) + `compute_pairwise_emd` (lidm/eval/metric_utils.py:447-458), `chamfer_3DDist` (reference lidm/eval/modules/chamfer3D/dist_chamfer_3D.py:28-76), `chamfer_2DDist`
(lidm/eval/modules/chamfer2D/dist_chamfer_2D.py) and `compute_pairwise_cd` (lidm/eval/metric_utils.py:414-423), on
the hand-written nearest-neighbour kernel behind `lidm_chamfer_nn`.  GPU tensors only, like the reference."""
import numpy as np
import torch

from . import _lib


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def _check_clouds(input1, input2, dim):
    for name, t in (("input1", input1), ("input2", input2)):
        if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 3 and t.shape[2] == dim):
            raise ValueError(f"{name} must be a CUDA fp32 tensor of shape (B, n, {dim})")
    if input1.shape[0] != input2.shape[0] or input1.device != input2.device:
        raise ValueError("input1 / input2 must share batch size and device")
    if min(input1.shape[0], input1.shape[1], input2.shape[1]) < 1:
        raise ValueError("empty point set")


class _ChamferFunction(torch.autograd.Function):
    """chamfer_3DFunction / chamfer_2DFunction (dist_chamfer_3D.py:28-66): forward = nearest neighbours both ways,
    backward = NmDistanceGradKernel (chamfer3D.cu:155-185)."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, dim, contract_fma):
        B, N, _ = xyz1.shape
        M = xyz2.shape[1]
        dev = xyz1.device
        dist1 = torch.empty((B, N), dtype=torch.float32, device=dev)
        dist2 = torch.empty((B, M), dtype=torch.float32, device=dev)
        idx1 = torch.empty((B, N), dtype=torch.int32, device=dev)
        idx2 = torch.empty((B, M), dtype=torch.int32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(lib.lidm_chamfer_nn_ex(xyz1.data_ptr(), xyz2.data_ptr(), B, N, M, dim, dist1.data_ptr(), idx1.data_ptr(),
                                              dist2.data_ptr(), idx2.data_ptr(), int(contract_fma), _stream_ptr(dev)))
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.dim = dim
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, graddist1, graddist2, gradidx1, gradidx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        graddist1, graddist2 = graddist1.contiguous(), graddist2.contiguous()
        gradxyz1, gradxyz2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
        B, N, _ = xyz1.shape
        lib = _lib.load()
        with torch.cuda.device(xyz1.device):
            _lib.check(lib.lidm_chamfer_backward(xyz1.data_ptr(), xyz2.data_ptr(), B, N, xyz2.shape[1], ctx.dim,
                                                 graddist1.data_ptr(), graddist2.data_ptr(), idx1.data_ptr(), idx2.data_ptr(),
                                                 gradxyz1.data_ptr(), gradxyz2.data_ptr(), _stream_ptr(xyz1.device)))
        return gradxyz1, gradxyz2, None, None


class _ChamferDist(torch.nn.Module):
    DIM = 3
    # how the squared distance is rounded: True = fma(dz,dz,fma(dx,dx,dy*dy)), the bits of the reference extension built by
    # nvcc (pinned on the GPU: tests/test_gpu_eval_ref.py); False = every operation rounded separately (the numpy form)
    CONTRACT_FMA = True

    def forward(self, input1, input2):
        """input1 (B,N,dim), input2 (B,M,dim) fp32 CUDA -> dist1 (B,N), dist2 (B,M) fp32, idx1 (B,N), idx2 (B,M) int32."""
        _check_clouds(input1, input2, self.DIM)
        return _ChamferFunction.apply(input1.contiguous(), input2.contiguous(), self.DIM, self.CONTRACT_FMA)


class chamfer_3DDist(_ChamferDist):
    DIM = 3


class chamfer_2DDist(_ChamferDist):
    DIM = 2


def compute_pairwise_cd(x, y, module=None):
    """reference lidm/eval/metric_utils.py:414-423: x (N,d) / y (M,d) numpy (or batched (B,*,d)) -> scalar Chamfer distance."""
    if x.ndim == 2 and y.ndim == 2:
        x, y = x[None], y[None]
    if module is None:
        module = chamfer_3DDist() if x.shape[-1] == 3 else chamfer_2DDist()
    xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
    yt = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).cuda()
    dist1, dist2, _, _ = module(xt, yt)
    return ((dist1.mean() + dist2.mean()) / 2).item()


def compute_pairwise_cd_batch(reference, samples):
    """reference lidm/eval/metric_utils.py:426-444: one reference cloud (Nr,d) against a list of sample clouds (Ns_i,d),
    d = 2 or 3.  Every cloud is padded to the longest one with far-away points (1e6), all pairs go through ONE kernel
    call, and each distance is averaged over the un-padded points only.  Returns a list of floats."""
    d = reference.ndim and reference.shape[-1]
    if reference.ndim != 2 or d not in (2, 3):
        raise ValueError("reference must be an (N, 2) or (N, 3) array")
    module = chamfer_3DDist() if d == 3 else chamfer_2DDist()
    len_r, len_s = reference.shape[0], [s.shape[0] for s in samples]
    width = max([len_r] + len_s)

    def padded(c):
        out = np.full((width, d), 1e6, dtype=np.float32)
        out[:c.shape[0]] = c
        return out

    ref = torch.from_numpy(padded(reference)).cuda()
    smp = torch.from_numpy(np.stack([padded(c) for c in samples])).cuda()
    dist_r, dist_s, _, _ = module(ref.expand_as(smp), smp)
    return [((dist_r[i, :len_r].mean() + dist_s[i, :n].mean()) / 2.).item() for i, n in enumerate(len_s)]


class emdFunction(torch.autograd.Function):
    """emdFunction (emd_module.py:47-91): the auction-algorithm EMD approximation; only xyz1 receives a gradient."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        batchsize, n, _ = xyz1.size()
        _, m, _ = xyz2.size()
        assert n == m
        assert xyz1.size()[0] == xyz2.size()[0]
        assert batchsize <= 512
        xyz1 = xyz1.contiguous().float().cuda()
        xyz2 = xyz2.contiguous().float().cuda()
        dist = torch.empty(batchsize, n, device=xyz1.device)
        assignment = torch.empty(batchsize, n, device=xyz1.device, dtype=torch.int32)
        lib = _lib.load()
        with torch.cuda.device(xyz1.device):
            _lib.check(lib.lidm_emd_forward(xyz1.data_ptr(), xyz2.data_ptr(), batchsize, n, float(eps), int(iters),
                                            dist.data_ptr(), assignment.data_ptr(), _stream_ptr(xyz1.device)))
        ctx.save_for_backward(xyz1, xyz2, assignment)
        ctx.mark_non_differentiable(assignment)
        return dist, assignment

    @staticmethod
    def backward(ctx, graddist, gradidx):
        xyz1, xyz2, assignment = ctx.saved_tensors
        graddist = graddist.contiguous()
        gradxyz1 = torch.empty_like(xyz1)
        lib = _lib.load()
        with torch.cuda.device(xyz1.device):
            _lib.check(lib.lidm_emd_backward(xyz1.data_ptr(), xyz2.data_ptr(), graddist.data_ptr(), assignment.data_ptr(),
                                             xyz1.shape[0], xyz1.shape[1], gradxyz1.data_ptr(), _stream_ptr(xyz1.device)))
        return gradxyz1, torch.zeros_like(xyz2), None, None


class emdModule(torch.nn.Module):
    """emd_module.py:94-98: forward(input1, input2, eps, iters) -> (dist (B,n) squared distances, assignment (B,n) int32)."""

    def forward(self, input1, input2, eps, iters):
        return emdFunction.apply(input1, input2, eps, iters)


def compute_pairwise_emd(x, y, module=None):
    """reference lidm/eval/metric_utils.py:447-458: clouds truncated to a multiple of 1024 points, eps 0.005, 50 iterations."""
    if module is None:
        module = emdModule()
    n_points = min(x.shape[0], y.shape[0])
    n_points = n_points - n_points % 1024
    x, y = x[:n_points], y[:n_points]
    if x.ndim == 2 and y.ndim == 2:
        x, y = x[None], y[None]
    x, y = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda(), torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)).cuda()
    dist, _ = module(x, y, 0.005, 50)
    return torch.sqrt(dist).mean().item()
