"""Multi-GPU plumbing for the extension path: read batches are sharded over ranks, there is no exchange step
(pairs are independent; SURVEY.md 8e, reference: single device, src/gasal2_ssw.cpp:34).  One process per GPU;
`torch.distributed` is used only for the barrier and for reducing the timing (max) and the work (sum).
Works with the `gloo` backend on CPU (tests) and `nccl` on GPUs (bench.py)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, disjoint, exhaustive split of [0, n_items) over `world` ranks (sizes differ by <= 1)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_seed(base_seed: int, rank: int) -> int:
    """Weak scaling: every rank aligns its own synthetic batch."""
    return base_seed + 1000003 * rank


def reduce_step(elapsed_ms: float, cells: float, dist=None, device="cpu") -> Tuple[float, float]:
    """(max over ranks of elapsed_ms, sum over ranks of cells)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(elapsed_ms), float(cells)
    import torch
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    c = torch.tensor([float(cells)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), float(c.item())
