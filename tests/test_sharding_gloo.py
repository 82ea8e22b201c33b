"""CPU tests of the multi-GPU host logic with world_size=2 over gloo (no GPU needed)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from htsp_b200.sharding import gather_tours, shard_range


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
    for total in (5, 8):
        port = _free_port()
        with mp.Manager() as mgr:
            ret = mgr.dict()
            mp.spawn(_worker, args=(2, port, total, 6, ret), nprocs=2, join=True)
            assert dict(ret) == {0: True, 1: True}
