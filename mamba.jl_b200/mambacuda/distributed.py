"""Cross-GPU diagnostics: the ONLY collective on the path.

Chains are sharded across ranks (one process per GPU, contiguous global chain ids); nothing is exchanged while sampling.
gelmandiag / summarystats over all chains run as the packed two-round protocol of libmambacuda (include/mambacuda.h,
csrc/diagproto.hpp): O(p) doubles per round.

  * A handle that joined an NCCL communicator (`init_comm`) does both rounds inside the library, on the device
    (`Engine.diag_global`): nothing in this module touches the data.
  * Any other transport carries the two buffers itself: `global_diagnostics` all-reduces them with torch.distributed (gloo in the
    CPU tests) around `local.diag_round1 / diag_round2 / diag_finish` — `local` is an Engine or anything with those methods.
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


def _allreduce(arr, op, device=None):
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return np.asarray(arr, dtype=np.float64)
    import torch
    t = torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64))
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}[op])
    return t.cpu().numpy()


def init_comm(eng):
    """Give `eng` the NCCL communicator of the torch.distributed world: rank 0 creates the id, torch broadcasts its 128 bytes."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return False
    from .engine import comm_unique_id
    box = [comm_unique_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(box, src=0)
    eng.comm_init(dist.get_rank(), dist.get_world_size(), box[0])
    return True


def global_diagnostics(local, alpha=0.05, transform=False, device=None):
    """(psrf [p x 2], summary [p x 5], link codes) over the chains of ALL ranks: src/output/gelmandiag.jl:3-60, stats.jl:85-94."""
    if hasattr(local, "comm_size") and local.comm_size()[1] > 1:
        return local.diag_global(alpha, transform)
    b1 = local.diag_round1()
    p = b1.size // 11
    r1 = np.concatenate([_allreduce(b1[:p], "min", device), _allreduce(b1[p:2 * p], "max", device), _allreduce(b1[2 * p:], "sum", device)])
    r2 = _allreduce(local.diag_round2(transform, r1), "sum", device)
    return local.diag_finish(alpha, transform, r1, r2)


def global_gelman(local, alpha=0.05, transform=False, device=None):
    return global_diagnostics(local, alpha, transform, device)[0]


def global_summary(local, device=None):
    return global_diagnostics(local, 0.05, False, device)[1]
