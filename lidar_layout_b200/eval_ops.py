"""Mirror of the reference's Chamfer-distance modules (forward only; the sampling / evaluation path never differentiates
through them): `chamfer_3DDist` (reference lidm/eval/modules/chamfer3D/dist_chamfer_3D.py:28-76), `chamfer_2DDist`
(lidm/eval/modules/chamfer2D/dist_chamfer_2D.py) and `compute_pairwise_cd` (lidm/eval/metric_utils.py:414-423), on
the hand-written nearest-neighbour kernel behind `lidm_chamfer_nn`.  GPU tensors only, like the reference."""
import numpy as np
import torch

from . import _lib


def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


class _ChamferDist(torch.nn.Module):
    DIM = 3

    def forward(self, input1, input2):
        """input1 (B,N,dim), input2 (B,M,dim) fp32 CUDA -> dist1 (B,N), dist2 (B,M) fp32, idx1 (B,N), idx2 (B,M) int32."""
        for name, t in (("input1", input1), ("input2", input2)):
            if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 3 and t.shape[2] == self.DIM):
                raise ValueError(f"{name} must be a CUDA fp32 tensor of shape (B, n, {self.DIM})")
        if input1.shape[0] != input2.shape[0] or input1.device != input2.device:
            raise ValueError("input1 / input2 must share batch size and device")
        input1, input2 = input1.contiguous(), input2.contiguous()
        B, N, _ = input1.shape
        M = input2.shape[1]
        if min(B, N, M) < 1:
            raise ValueError("empty point set")
        dev = input1.device
        dist1 = torch.empty((B, N), dtype=torch.float32, device=dev)
        dist2 = torch.empty((B, M), dtype=torch.float32, device=dev)
        idx1 = torch.empty((B, N), dtype=torch.int32, device=dev)
        idx2 = torch.empty((B, M), dtype=torch.int32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(lib.lidm_chamfer_nn(input1.data_ptr(), input2.data_ptr(), B, N, M, self.DIM, dist1.data_ptr(),
                                           idx1.data_ptr(), dist2.data_ptr(), idx2.data_ptr(), _stream_ptr(dev)))
        return dist1, dist2, idx1, idx2


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
