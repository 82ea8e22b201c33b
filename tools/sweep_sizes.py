#!/usr/bin/env python
"""BASELINE config 5: subproblem-size sweep N = G in {100, 200, 500} x batch 1K ... 64K at 1 / 2 / 4 / 8 GPUs
(device-resident inputs, noAug_nTraj, greedy, L = 6).  One JSON line per (N, batch): whole-job solves/s, ms per
step, which rollout kernel served it (tcgen05 for N <= 208 and, in mode fast, in key blocks up to N = 576; mma.sync above) and the algorithmic TFLOP/s of the whole
step (SURVEY 8d count).  Under torchrun (one rank per GPU) the batch is PER GPU (weak scaling, as bench.py): every
rank solves its own subproblems, the selected tours are all-gathered over NCCL each step, the time is the maximum
over ranks (CUDA events between barriers).

  python tools/sweep_sizes.py [--sizes 100,200,500] [--steps 2] [--mode fast]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_sizes.py"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import htsp_b200  # noqa: E402,F401
from htsp_b200 import B200PathSolver  # noqa: E402
from htsp_b200.weights import synthetic_state_dict  # noqa: E402

BATCHES = {100: [1024, 4096, 16384, 65536], 200: [1024, 4096, 16384, 65536], 500: [1024, 4096]}


def flops(N, G, L, H=128):
    enc = L * (24 * N * H * H + 4 * N * N * H) + 4 * N * H
    reset = 2 * (3 * N * H * H + (G + 3) * H * H)
    dec = (N - 2) * G * (6 * N * H + 2 * H * H)
    return enc + reset + dec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="100,200,500")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--mode", default="fast")
    ap.add_argument("--max-batch", type=int, default=65536)
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        from htsp_b200.sharding import gather_tours
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    sd = synthetic_state_dict(1234, 6)
    for N in [int(s) for s in a.sizes.split(",")]:
        model = B200PathSolver(state_dict=sd, mode=a.mode, device=dev, graph_size=N)
        model.first_action_override = torch.randperm(N, generator=torch.Generator().manual_seed(7))[:N]
        for B in BATCHES.get(N, [1024, 4096]):
            if B > a.max_batch:
                continue
            g = torch.Generator().manual_seed(1234 + N + 97 * rank)
            xs = [torch.rand(B, N, 2, generator=g).to(dev) for _ in range(2)]
            src = torch.zeros(B, 1, dtype=torch.long, device=dev)
            tgt = torch.full((B, 1), N - 1, dtype=torch.long, device=dev)
            length, pi = model(xs[0], src, tgt, val_type="noAug_nTraj", return_pi=True, group_size=N)      # warm-up
            if world > 1:
                gather_tours(pi, length, B * world)
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.steps):
                length, pi = model(xs[(i + 1) % 2], src, tgt, val_type="noAug_nTraj", return_pi=True, group_size=N)
                if world > 1:
                    gpi, glen = gather_tours(pi, length, B * world)
            e1.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t_ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
            okt = torch.tensor([int(bool((pi.sort(dim=-1).values == torch.arange(N, device=dev)).all()))], device=dev)
            if world > 1:
                dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
                dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            ms, ok = float(t_ms), bool(int(okt))
            if rank != 0:
                del xs, length, pi
                torch.cuda.empty_cache()
                continue
            B_total = B * world
            print(json.dumps({"metric": f"{N}-node subpath solves/sec (POMO {N} starts, greedy)", "nodes": N, "batch_per_gpu": B,
                              "n_gpus": world, "scaling": "weak",
                              "value": B_total / (ms * 1e-3), "unit": "solves/s", "ms_per_step": ms, "arith": a.mode,
                              "rollout_kernel": "tcgen05" if (N <= 208 or (a.mode == "fast" and N <= 576)) else "mma.sync",
                              "algorithmic_tflops": flops(N, N, 6) * B_total / (ms * 1e-3) / 1e12,
                              "tours_are_permutations": ok, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}),
                  flush=True)
            del xs, length, pi
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()


    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
