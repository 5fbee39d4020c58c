"""Multi-GPU host logic: alignment columns are independent (SubRecon.java:112-116: one task per column, no
shared mutable state), so contiguous column blocks are sharded across ranks — one process per GPU — with NO
data-path collective. The only exchanges are a gather of per-column results to rank 0 (what the ordered
`Future.get()` join does) and a reduce of the total log-likelihood (SubRecon.java:137,143).
"""
from __future__ import annotations

import numpy as np


def column_block(n_sites: int, world_size: int, rank: int):
    """Contiguous block of rank `rank`: [g*ceil(L/G), min(L,(g+1)*ceil(L/G))). Returns (site0, n)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    per = -(-n_sites // world_size)
    lo = min(n_sites, rank * per)
    hi = min(n_sites, lo + per)
    return lo, hi - lo


def gather_lnl(local_lnl: np.ndarray, n_sites: int, group=None):
    """All ranks call this with their block's lnl; rank 0 gets lnl[n_sites] in column order, others None."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-n_sites // world)
    buf = torch.zeros(per, dtype=torch.float64)
    buf[:len(local_lnl)] = torch.from_numpy(np.ascontiguousarray(local_lnl))
    if dist.get_backend(group) == "nccl":
        buf = buf.cuda()
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, outs, dst=0, group=group)
    if rank != 0:
        return None
    full = torch.cat([o.cpu() for o in outs]).numpy()[:n_sites]
    return full


def reduce_total_lnl(local_lnl: np.ndarray, group=None) -> float:
    """Sum of the per-column lnL over all ranks (an all-reduce of ONE double per rank). The reference adds
    columns sequentially on the host (SubRecon.java:137); the tree-order sum here agrees to rounding and is
    used as a cross-check next to the ordered host sum of `gather_lnl`."""
    import torch
    import torch.distributed as dist
    s = 0.0
    for v in local_lnl:
        s += float(v)
    t = torch.tensor([s], dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())
