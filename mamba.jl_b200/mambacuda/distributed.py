"""Cross-GPU diagnostics: the ONLY collective on the path.

Chains are sharded across ranks (one process per GPU, contiguous global chain ids); nothing is
exchanged while sampling.  gelmandiag / summarystats over all chains need cross-chain sums of
per-chain moments: each rank reduces its own chains on the device (mcu_moments / mcu_summary_sums),
and the O(p) partial sums are all-reduced with torch.distributed (NCCL over NVLink on GPUs, gloo in
the CPU tests).  `local` is anything that offers the Engine's reduction methods.
"""
import numpy as np


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def _allreduce(arr, op="sum", device=None):
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return np.asarray(arr, dtype=np.float64)
    import torch
    t = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64))
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    rop = {"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}[op]
    dist.all_reduce(t, op=rop)
    return t.cpu().numpy()


def global_link_codes(local, transform, device=None):
    mm = local.minmax()
    mm = np.stack([_allreduce(mm[:, 0], "min", device), _allreduce(mm[:, 1], "max", device)], axis=1)
    return local.link_codes(transform, mm)


def global_gelman(local, alpha=0.05, transform=False, device=None):
    """gelmandiag over the chains of ALL ranks (src/output/gelmandiag.jl:3-60): two all-reduces of 7p doubles."""
    codes = global_link_codes(local, transform, device) if transform else None
    s0, n = local.moments(codes, None)
    s0 = _allreduce(s0, "sum", device)
    center = np.stack([s0[:, 1] / s0[:, 0], s0[:, 3] / s0[:, 0]], axis=1)
    s1, n = local.moments(codes, center)
    s1 = _allreduce(s1, "sum", device)
    return local.gelman_from_moments(n, center, s1, alpha)


def global_summary(local, device=None):
    """Streaming summarystats over the chains of all ranks (src/output/stats.jl:85-94): [p × 5]."""
    s0 = _allreduce(local.summary_sums(None), "sum", device)
    nb = np.where(s0[:, 4] > 0, s0[:, 4], 1.0)
    center = np.stack([s0[:, 1] / s0[:, 0], s0[:, 5] / nb], axis=1)
    s1 = _allreduce(local.summary_sums(center), "sum", device)
    n = local.moments(None, None)[1]
    return local.summary_from_sums(n, center, s1)
