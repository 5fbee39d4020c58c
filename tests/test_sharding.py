"""Multi-GPU host logic on CPU: world_size-2 gloo run of the column sharding + gather + total-lnL reduce.
The per-rank compute is the oracle here (tests may use it); on GPUs it is one `native.Context` per rank."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from subrecon_b200 import sharding


def test_column_blocks_cover_everything():
    for L in (0, 1, 7, 130, 1000, 1_000_003):
        for G in (1, 2, 3, 4, 8):
            blocks = [sharding.column_block(L, G, r) for r in range(G)]
            assert sum(n for _, n in blocks) == L
            pos = 0
            for lo, n in blocks:
                assert lo == pos or n == 0
                pos += n
    with pytest.raises(ValueError):
        sharding.column_block(10, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import load_example
        from oracle import oracle
        p = load_example()
        L = p.states.shape[1]
        lo, n = sharding.column_block(L, world, rank)
        lnl, _ = oracle.run(p.flat, p.states, p.model, p.rates, site0=lo, n=n, want_joint=False)
        full = sharding.gather_lnl(lnl, L)
        tot = sharding.reduce_total_lnl(lnl)
        if rank == 0:
            q.put((full, tot))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_single_run(example_problem):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    full, tot = q.get(timeout=120)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    from oracle import oracle
    p = example_problem
    want, _ = oracle.run(p.flat, p.states, p.model, p.rates, want_joint=False)
    assert np.array_equal(full, want)                       # gathered in column order, bit-identical
    assert abs(tot - want.sum()) < 1e-9
