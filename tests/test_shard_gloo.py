"""N > 1 host logic on CPU: two `gloo` ranks shard the chains, run their blocks (with the
oracle standing in for the GPU kernels -- allowed in tests/) and gather on rank 0.  The
gathered draws must equal the single-process run chain by chain (world-size invariance)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from bsreg_b200.shard import chain_block, fit_sharded
from bsreg_b200.synthetic import make_problem
from oracle.sampler import PhiloxRNG, RefChain, sample

N_CHAINS, N_SAVE, N_BURN, SEED = 5, 12, 8, 77


def test_chain_blocks_partition_the_chains():
    for n in (1, 5, 64, 512):
        for world in (1, 2, 3, 4, 8):
            if world > n:
                with pytest.raises(ValueError):
                    chain_block(n, 0, world)
                continue
            blocks = [chain_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            assert all(blocks[r][0] + blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    assert chain_block(512, 3, 8) == (192, 64) and chain_block(5, 1, 2) == (3, 2)


def _run_block(lo, cnt):
    p = make_problem("sar", n=150, k_reg=3, knn=4, symmetric=True)
    return np.stack([sample(RefChain("sar", p.y, p.X, W=p.W, rng=PhiloxRNG(SEED, lo + c)), N_SAVE, N_BURN)
                     for c in range(cnt)])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = fit_sharded(_run_block, N_CHAINS)
        if rank == 0:
            q.put(out)
        else:
            assert out is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_fit_equals_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    want = _run_block(0, N_CHAINS)
    assert got.shape == (N_CHAINS, N_SAVE, 5)
    assert np.array_equal(got, want)
