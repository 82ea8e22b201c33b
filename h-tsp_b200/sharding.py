"""Multi-GPU plumbing: subproblems are independent, so a batch is split
into contiguous blocks over the ranks of one node (one process per GPU), the 5 MB of weights are
replicated, and the only communication is one all-gather of the selected tours + lengths
(N*8+4 bytes per subproblem) over NCCL/NVLink.  No data-path collective exists in the rollout.
When there are fewer subproblems than ranks (one TSP-10000 instance => B = 1 per environment step)
the 8 augmentations of the hook are sharded instead (`solve_aug_sharded`): every rank rolls out its
augmentations of every subproblem and the per-augmentation winners ((N ids + 1 float) x 8 x B) are
all-gathered and reduced with the reference's first-max rule (train_path_solver.py:297-306)."""
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
    # a failed shard must not leave the other ranks waiting in the collective: the status travels with the lengths
    # (NaN) and every rank raises after the all-gather
    lengths = torch.where(status == 0, lengths, torch.full_like(lengths, float("nan")))
    all_paths, all_len = gather_tours(paths, lengths, B)
    bad = torch.isnan(all_len)
    if bool(bad.any()):
        raise RuntimeError(f"invalid path(s) or out-of-range fragment ids in subproblem(s) "
                           f"{torch.nonzero(bad).reshape(-1).tolist()} (reported on every rank)")
    return all_paths, all_len


# ---- B < ranks: shard the 8 augmentations ---------------------------------
@torch.no_grad()
def solve_aug_shard(solver, data: torch.Tensor, fragment: torch.Tensor, a_lo: int, a_hi: int,
                    first: torch.Tensor):
    """Rolls out augmentations [a_lo, a_hi) of EVERY subproblem of the hook call on this rank's GPU.
    Returns (best_len [A_r,B] f32, best_pi [A_r,B,N] i64 local ids, scale [B], fragment on device):
    the winner over the G POMO starts of each (augmentation, subproblem)."""
    from . import _abi
    lib = _abi.lib()
    m = solver._solver_model
    dev = m.device
    data = data.to(dev, torch.float32).contiguous()
    fragment = fragment.to(dev, torch.int64).contiguous()
    B, Ntot, _ = data.shape
    N = fragment.shape[1]
    x_new = torch.empty(B, N, 2, dtype=torch.float32, device=dev)
    scale = torch.empty(B, dtype=torch.float32, device=dev)
    x_aug = torch.empty(8 * B, N, 2, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _abi.check(lib.htsp_extract_normalize(data.data_ptr(), fragment.data_ptr(), B, Ntot, N, 8,
                                              x_new.data_ptr(), scale.data_ptr(), x_aug.data_ptr(),
                                              m._stream()), "htsp_extract_normalize")
    A = a_hi - a_lo
    if A == 0:
        return (torch.empty(0, B, dtype=torch.float32, device=dev),
                torch.empty(0, B, N, dtype=torch.int64, device=dev), scale, fragment)
    xs = x_aug[a_lo * B:a_hi * B]                          # augmentations are block-major (utils.py:54)
    src = torch.zeros(A * B, dtype=torch.int32, device=dev)
    tgt = torch.full((A * B,), N - 1, dtype=torch.int32, device=dev)
    pi, reward = m._rollout_chunked(xs, src, tgt, first.to(dev), 1)
    best_len, best_pi, _ = m.select_best(reward, pi, 1)     # max over the G starts, first max wins
    return best_len.reshape(A, B), best_pi.reshape(A, B, N), scale, fragment


def combine_aug_winners(best_len: torch.Tensor, best_pi: torch.Tensor):
    """[8,B] lengths and [8,B,N] tours in augmentation order -> the reference's reduction over the
    augmentation axis: maximum reward = minimum length, the FIRST augmentation wins ties
    (`aug_reward.max(dim=0)` on the [8,B] matrix, train_path_solver.py:300-306)."""
    A, B = best_len.shape
    win_len = best_len[0].clone()
    win_pi = best_pi[0].clone()
    for a in range(1, A):
        better = best_len[a] < win_len                      # strict: an equal later augmentation never wins
        win_len = torch.where(better, best_len[a], win_len)
        win_pi = torch.where(better[:, None], best_pi[a], win_pi)
    return win_len, win_pi


def gather_aug_winners(best_len: torch.Tensor, best_pi: torch.Tensor, world: int, rank: int):
    """All-gathers the per-rank augmentation winners ([A_r,B], [A_r,B,N]) into augmentation order
    ([8,B], [8,B,N]) on every rank; shards are padded to the largest one for the collective."""
    if world == 1:
        return best_len, best_pi
    B, N = best_pi.shape[1], best_pi.shape[2]
    cap = -(-8 // world)
    pad_len = best_len.new_zeros(cap, B)
    pad_pi = best_pi.new_zeros(cap, B, N)
    pad_len[: best_len.shape[0]] = best_len
    pad_pi[: best_pi.shape[0]] = best_pi
    all_len = best_len.new_empty(world * cap, B)
    all_pi = best_pi.new_empty(world * cap, B, N)
    dist.all_gather_into_tensor(all_len, pad_len)
    dist.all_gather_into_tensor(all_pi, pad_pi)
    keep = []
    for r in range(world):
        lo, hi = shard_range(8, r, world)
        keep.append(torch.arange(r * cap, r * cap + (hi - lo), device=best_pi.device))
    idx = torch.cat(keep)
    return all_len.index_select(0, idx), all_pi.index_select(0, idx)


@torch.no_grad()
def solve_aug_sharded(solver, data: torch.Tensor, fragment: torch.Tensor):
    """Hook call with fewer subproblems than ranks: every rank holds (data, fragment), solves its
    share of the 8 augmentations and all ranks end with (paths [B,N] i64 global ids, lengths [B]).
    The POMO starts are drawn on rank 0 and broadcast so that the result equals the one-GPU call."""
    from . import _abi
    lib = _abi.lib()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    m = solver._solver_model
    N = fragment.shape[1]
    first = m.draw_first_action(N, solver._sample_size).to(m.device).contiguous()
    if world > 1:
        dist.broadcast(first, src=0)
    a_lo, a_hi = shard_range(8, rank, world)
    best_len, best_pi, scale, frag_dev = solve_aug_shard(solver, data, fragment, a_lo, a_hi, first)
    all_len, all_pi = gather_aug_winners(best_len, best_pi, world, rank)
    win_len, win_pi = combine_aug_winners(all_len, all_pi)
    B = win_len.shape[0]
    dev = m.device
    paths = torch.empty(B, N, dtype=torch.int64, device=dev)
    lengths = torch.empty(B, dtype=torch.float32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _abi.check(lib.htsp_finalize_paths(win_pi.contiguous().data_ptr(), frag_dev.data_ptr(),
                                           win_len.contiguous().data_ptr(), scale.data_ptr(), B, N,
                                           paths.data_ptr(), lengths.data_ptr(), status.data_ptr(),
                                           m._stream()), "htsp_finalize_paths")
    assert bool((status == 0).all())
    return paths, lengths
