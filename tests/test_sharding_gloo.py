"""CPU tests of the multi-GPU host logic with world_size=2 over gloo (no GPU needed)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from htsp_b200.sharding import combine_aug_winners, gather_aug_winners, gather_tours, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 16, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, N, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(total, rank, world)
        full_pi = torch.arange(total * N, dtype=torch.int64).reshape(total, N)
        full_len = torch.arange(total, dtype=torch.float32) * 0.5
        pi, ln = gather_tours(full_pi[lo:hi].clone(), full_len[lo:hi].clone(), total)
        ret[rank] = bool(torch.equal(pi, full_pi) and torch.equal(ln, full_len))
    finally:
        dist.destroy_process_group()


def test_gather_tours_world2_ragged():
    for total in (1, 5, 8):          # total = 1: rank 1 holds an empty shard (fewer subproblems than ranks)
        port = _free_port()
        with mp.Manager() as mgr:
            ret = mgr.dict()
            mp.spawn(_worker, args=(2, port, total, 6, ret), nprocs=2, join=True)
            assert dict(ret) == {0: True, 1: True}


def test_combine_aug_winners_first_min_semantics():
    """The reduction over the augmentation axis must be `aug_reward.max(dim=0)` of the reference
    (train_path_solver.py:300-306): smallest length, FIRST augmentation on ties."""
    g = torch.Generator().manual_seed(0)
    ln = torch.randint(1, 4, (8, 50), generator=g).float()          # many exact ties
    pi = torch.arange(8 * 50 * 5).reshape(8, 50, 5)
    win_len, win_pi = combine_aug_winners(ln, pi)
    want_r, want_a = (-ln).max(dim=0)                               # CPU max: first index among equals
    assert torch.equal(win_len, -want_r)
    assert torch.equal(win_pi, pi[want_a, torch.arange(50)])


def _aug_worker(rank, world, port, B, N, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        full_len = torch.rand(8, B, generator=g)
        full_pi = torch.randint(0, N, (8, B, N), generator=g)
        lo, hi = shard_range(8, rank, world)
        ln, pi = gather_aug_winners(full_len[lo:hi].clone(), full_pi[lo:hi].clone(), world, rank)
        ret[rank] = bool(torch.equal(ln, full_len) and torch.equal(pi, full_pi))
    finally:
        dist.destroy_process_group()


def test_gather_aug_winners_world2_and_3():
    for world in (2, 3):       # 8 augmentations over 3 ranks: ragged 3/3/2
        port = _free_port()
        with mp.Manager() as mgr:
            ret = mgr.dict()
            mp.spawn(_aug_worker, args=(world, port, 2, 7, ret), nprocs=world, join=True)
            assert dict(ret) == {r: True for r in range(world)}
