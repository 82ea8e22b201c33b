"""Multi-GPU plumbing: subproblems are independent, so a batch is split
into contiguous blocks over the ranks of one node (one process per GPU), the 5 MB of weights are
replicated, and the only communication is one all-gather of the selected tours + lengths
(N*8+4 bytes per subproblem) over NCCL/NVLink.  No data-path collective exists in the rollout."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `total` items owned by `rank`; sizes differ by at most 1."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_tours(best_pi: torch.Tensor, lengths: torch.Tensor, total: int):
    """All-gathers per-rank (best_pi [b_r,N] i64, lengths [b_r] f32) into ([total,N], [total]) on
    every rank.  Ragged shards are padded to the largest shard for the collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return best_pi, lengths
    world, rank = dist.get_world_size(), dist.get_rank()
    N = best_pi.shape[1]
    cap = -(-total // world)
    pad_pi = best_pi.new_zeros(cap, N)
    pad_len = lengths.new_zeros(cap)
    pad_pi[: best_pi.shape[0]] = best_pi
    pad_len[: lengths.shape[0]] = lengths
    all_pi = best_pi.new_empty(world * cap, N)
    all_len = lengths.new_empty(world * cap)
    dist.all_gather_into_tensor(all_pi, pad_pi)
    dist.all_gather_into_tensor(all_len, pad_len)
    keep = []
    for r in range(world):
        lo, hi = shard_range(total, r, world)
        keep.append(torch.arange(r * cap, r * cap + (hi - lo), device=best_pi.device))
    idx = torch.cat(keep)
    return all_pi.index_select(0, idx), all_len.index_select(0, idx)


@torch.no_grad()
def solve_sharded(solver, data: torch.Tensor, fragment: torch.Tensor):
    """Hook-level data parallelism: every rank holds the full (data, fragment) batch on the host,
    solves its contiguous shard on its own GPU with `solver.solve_device` and all ranks end with
    the complete (paths [B,N] i64, lengths [B] f32)."""
    B = fragment.shape[0]
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(B, rank, world)
    paths, lengths, status, _ = solver.solve_device(data[lo:hi], fragment[lo:hi], None)
    assert bool((status == 0).all())
    return gather_tours(paths, lengths, B)
