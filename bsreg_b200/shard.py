"""Chain sharding across the GPUs of one box (one process per GPU, torch.distributed).

Chains are the independent units of the Gibbs sampler, so the shard is a contiguous block of
global chain ids per rank and the inner loop has NO collective: W, y, X are replicated, every
rank runs its block with Philox keyed by the GLOBAL chain id (draws are therefore identical
for any world size), and the draws are gathered to rank 0 once at the end.
The same split is used inside one process by `bsreg_options.devices` (csrc/api.cu).
"""
from __future__ import annotations

import numpy as np


def chain_block(n_chains: int, rank: int, world: int) -> tuple[int, int]:
    """(first global chain id, number of chains) of `rank`: contiguous blocks, the first
    `n_chains % world` ranks carry one extra chain (same rule as bsreg_create)."""
    if not 0 <= rank < world:
        raise ValueError("rank outside [0, world)")
    if world > n_chains:
        raise ValueError("more ranks than chains")
    base, extra = divmod(n_chains, world)
    lo = rank * base + min(rank, extra)
    return lo, base + (1 if rank < extra else 0)


def gather_draws(local: np.ndarray, n_chains: int, dst: int = 0, device=None):
    """Gather per-rank draws (c_local, n_save, P) into (n_chains, n_save, P) on `dst`, in global
    chain order.  One collective per fit, after the sampling loop; returns None on other ranks."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local)
    rank, world = dist.get_rank(), dist.get_world_size()
    _, n_save, P = local.shape
    counts = [chain_block(n_chains, r, world)[1] for r in range(world)]
    cmax = max(counts)
    buf = torch.zeros((cmax, n_save, P), dtype=torch.float64, device=device)
    buf[: local.shape[0]] = torch.as_tensor(np.ascontiguousarray(local), device=device)
    outs = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, outs, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(outs, counts)], axis=0)


def fit_sharded(run_block, n_chains: int, device=None):
    """Run `run_block(chain_lo, count) -> (count, n_save, P)` on this rank's block of chains and
    gather on rank 0.  `run_block` is typically a closure over `interface.get_model(...,
    n_chains=count, chain_offset=lo).sample(...)` bound to this rank's GPU."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, cnt = chain_block(n_chains, rank, world)
    return gather_draws(run_block(lo, cnt), n_chains, device=device)
