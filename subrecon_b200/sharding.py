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


def run_multi_context(contexts, states, want_joint=True, want_compact=False, threshold=0.5):
    """Single-process, multi-GPU form of the sharding (what a JVM host does: one `sr_ctx` per device, each driven by its
    own thread). `contexts` are `native.Context`s with model, rates and tree already set; columns are
    split with `column_block`, every context streams its block from host memory, results are concatenated in column
    order. ctypes releases the GIL during the native call, so the blocks really run concurrently."""
    import threading
    states = np.ascontiguousarray(states, dtype=np.uint8)
    n_sites = states.shape[1]
    G = len(contexts)
    outs = [None] * G
    errs = [None] * G

    def work(g):
        try:
            lo, n = column_block(n_sites, G, g)
            outs[g] = contexts[g].run_streamed(states, lo, n, want_joint=want_joint, want_compact=want_compact,
                                               threshold=threshold) if n else None
        except Exception as e:          # surfaced on the caller's thread
            errs[g] = e

    threads = [threading.Thread(target=work, args=(g,)) for g in range(G)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    parts = [o for o in outs if o is not None]
    res = {"lnl": np.concatenate([o["lnl"] for o in parts]) if parts else np.empty(0)}
    res["joint"] = np.concatenate([o["joint"] for o in parts]) if want_joint and parts else None
    res["compact"] = np.concatenate([o["compact"] for o in parts]) if want_compact and parts else None
    return res
