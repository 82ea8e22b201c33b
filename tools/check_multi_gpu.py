#!/usr/bin/env python
"""Multi-GPU check on real devices, launched with torchrun, one rank per GPU over NCCL:
  (1) batch sharding: `solve_sharded` over R ranks returns exactly the tours / lengths of the one-GPU call;
  (2) augmentation sharding (B < R, e.g. one TSP-10000 instance): `solve_aug_sharded` == one-GPU x8Aug hook call.
Rank 0 prints one JSON line.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/check_multi_gpu.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import htsp_b200  # noqa: E402,F401
from htsp_b200 import B200PathSolver, B200RLSolver, sharding  # noqa: E402
from htsp_b200.weights import synthetic_state_dict  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    model = B200PathSolver(state_dict=synthetic_state_dict(1234, 6), mode="x3", device=dev)
    hook = B200RLSolver(model, sample_size=200)
    g = torch.Generator().manual_seed(5)
    Ntot, N = 2000, 200
    out = {"world": world}
    for name, B in (("batch_sharded", 4 * world + 1), ("aug_sharded", 1)):
        data = torch.rand(B, Ntot, 2, generator=g).to(dev)
        frag = torch.stack([torch.randperm(Ntot, generator=g)[:N] for _ in range(B)]).to(dev)
        first = torch.randperm(N, generator=torch.Generator().manual_seed(3))[:200]
        model.first_action_override = first
        ref_paths, ref_len, _, _ = hook.solve_device(data, frag, None)            # every rank: the one-GPU answer
        if name == "batch_sharded":
            paths, lens = sharding.solve_sharded(hook, data, frag)
        else:
            paths, lens = sharding.solve_aug_sharded(hook, data, frag)
        same = bool(torch.equal(paths.to(dev), ref_paths)) and bool(torch.allclose(lens.to(dev), ref_len, rtol=0, atol=0))
        flag = torch.tensor([int(same)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out[name] = {"subproblems": B, "identical_to_one_gpu": bool(flag.item())}
    if rank == 0:
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
