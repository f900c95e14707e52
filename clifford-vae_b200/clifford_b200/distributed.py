"""Multi-GPU host logic: one process per GPU (torch.distributed, NCCL over NVLink on the box, gloo in
the CPU tests).  Every latent / VSA op is independent per row, so rows are sharded
with NO collective on the data path; collectives appear only where the reference's math reduces over
rows: `bundle` over a row-sharded stack (all-reduce of d floats) and cleanup similarity against a
row-sharded item memory ((max, argmax) all-gather).  The local compute defaults to the CUDA kernels;
the `local_*` hooks exist so the collective logic can be tested on CPU with gloo.
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_rows(total: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> Tuple[int, int]:
    """[start, stop) of this rank's contiguous, balanced slice of `total` rows (first ranks get the remainder)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, rem = divmod(total, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def rank_seed(seed: int, rank: Optional[int] = None) -> int:
    """Rank-disjoint Philox seed (the same mixing _lib.next_rng applies under an initialised group)."""
    rank = world()[0] if rank is None else rank
    return (seed ^ (rank * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF


def max_over_ranks(value: float, device=None) -> float:
    """Timing convention of bench.py: the slowest rank defines the step."""
    if world()[1] == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sharded_bundle(local_vectors: torch.Tensor, k_total: int, normalize: bool = True, group=None,
                   local_sum: Optional[Callable[[torch.Tensor], torch.Tensor]] = None) -> torch.Tensor:
    """bundle (reference utils/vsa.py:75-79) of a stack whose k_total rows are sharded over ranks:
    local partial sum (bundle kernel), all-reduce(sum) of d floats, then / sqrt(k_total)."""
    if local_sum is None:
        from . import vsa
        local_sum = lambda v: vsa.bundle(v, normalize=False)   # noqa: E731
    part = local_sum(local_vectors).contiguous()
    if world()[1] > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    return part / math.sqrt(k_total) if normalize else part


def sharded_cleanup(query: torch.Tensor, local_items: torch.Tensor, global_offset: int, group=None,
                    local_similarity: Optional[Callable[[torch.Tensor, torch.Tensor], torch.Tensor]] = None):
    """argmax_j cos(query_q, item_j) over an item memory sharded by rows (reference cleanup step
    utils/vsa.py:316-317, scripts/rolefiller_heatmap.py:17-44).  query (Q, d) replicated; local_items
    (M_local, d) with global row index global_offset + local index.  Returns (best_sim (Q,), best_idx (Q,))."""
    if local_similarity is None:
        from . import vsa
        # many-query cleanup is a (Q x d) x (d x M) GEMM on unit vectors (cuBLAS), as SURVEY 8(d) allows
        local_similarity = lambda q, m: vsa.normalize_vectors(q) @ vsa.normalize_vectors(m).T   # noqa: E731
    sims = local_similarity(query, local_items)              # (Q, M_local)
    best, arg = sims.max(dim=1)
    arg = arg + global_offset
    rank, w = world()
    if w == 1:
        return best, arg
    bests = [torch.empty_like(best) for _ in range(w)]
    args = [torch.empty_like(arg) for _ in range(w)]
    dist.all_gather(bests, best.contiguous(), group=group)
    dist.all_gather(args, arg.contiguous(), group=group)
    bests, args = torch.stack(bests), torch.stack(args)      # (W, Q)
    win = bests.argmax(dim=0)                                  # first (lowest-rank) maximum, like a global argmax
    q = torch.arange(bests.shape[1], device=bests.device)
    return bests[win, q], args[win, q]
