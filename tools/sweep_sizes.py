#!/usr/bin/env python
"""BASELINE config 5: subproblem-size sweep N = G in {100, 200, 500} x batch 1K ... 64K on one GPU
(device-resident inputs, noAug_nTraj, greedy, L = 6, x3 arithmetic).  One JSON line per (N, batch):
solves/s, ms per step, which rollout kernel served it (tcgen05 for N <= 208, mma.sync above) and the
algorithmic TFLOP/s of the whole step (SURVEY 8d count).  Multi-GPU scaling is bench.py's job
(`--gpus N --nodes .. --batch ..` runs the same step sharded over ranks).

  python tools/sweep_sizes.py [--sizes 100,200,500] [--steps 2]"""
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
    ap.add_argument("--mode", default="x3")
    ap.add_argument("--max-batch", type=int, default=65536)
    a = ap.parse_args()
    dev = torch.device("cuda")
    sd = synthetic_state_dict(1234, 6)
    for N in [int(s) for s in a.sizes.split(",")]:
        model = B200PathSolver(state_dict=sd, mode=a.mode, device=dev, graph_size=N)
        model.first_action_override = torch.randperm(N, generator=torch.Generator().manual_seed(7))[:N]
        for B in BATCHES.get(N, [1024, 4096]):
            if B > a.max_batch:
                continue
            g = torch.Generator().manual_seed(1234 + N)
            xs = [torch.rand(B, N, 2, generator=g).to(dev) for _ in range(2)]
            src = torch.zeros(B, 1, dtype=torch.long, device=dev)
            tgt = torch.full((B, 1), N - 1, dtype=torch.long, device=dev)
            model(xs[0], src, tgt, val_type="noAug_nTraj", return_pi=True, group_size=N)      # warm-up
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.steps):
                length, pi = model(xs[(i + 1) % 2], src, tgt, val_type="noAug_nTraj", return_pi=True, group_size=N)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            ok = bool((pi.sort(dim=-1).values == torch.arange(N, device=dev)).all())
            print(json.dumps({"metric": f"{N}-node subpath solves/sec (POMO {N} starts, greedy)", "nodes": N, "batch": B,
                              "value": B / (ms * 1e-3), "unit": "solves/s", "ms_per_step": ms, "arith": a.mode,
                              "rollout_kernel": "tcgen05" if N <= 208 else "mma.sync",
                              "algorithmic_tflops": flops(N, N, 6) * B / (ms * 1e-3) / 1e12,
                              "tours_are_permutations": ok, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}),
                  flush=True)
            del xs, length, pi
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
