"""Measurement utility (not a test): eager-PyTorch GPU throughput of the oracle port, i.e. the
reference's op sequence (ATen/cuBLAS kernels, fp32, TF32 off) on the same B200, as context for the
">20x the reference's single-GPU PyTorch throughput" target.  The port has no per-step host syncs
and no O(t) list growth, so it is an UPPER bound on what the unmodified reference reaches.
    python tests/baseline_torch_gpu.py [--batch 256] [--tf32]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import htsp_b200  # noqa: E402,F401
from htsp_b200.weights import synthetic_state_dict  # noqa: E402
from oracle import path_solver_oracle as po  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--nodes", type=int, default=200)
    ap.add_argument("--group", type=int, default=200)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--tf32", action="store_true")
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = a.tf32
    torch.backends.cudnn.allow_tf32 = a.tf32
    dev = torch.device("cuda")
    sd = {k: v.to(dev) for k, v in synthetic_state_dict(1234, 6).items()}
    g = torch.Generator().manual_seed(1)
    x = torch.rand(a.batch, a.nodes, 2, generator=g).to(dev)
    src = torch.zeros(a.batch, 1, dtype=torch.long, device=dev)
    tgt = torch.full((a.batch, 1), a.nodes - 1, dtype=torch.long, device=dev)
    first = torch.randperm(a.nodes, generator=g)[:a.group].to(dev)
    best = None
    with torch.no_grad():
        for _ in range(a.reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            po.path_solver_forward(sd, x, src, tgt, first, val_type="noAug_nTraj", return_pi=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    print(json.dumps({"baseline": "oracle port, eager PyTorch on GPU", "tf32": a.tf32,
                      "batch": a.batch, "nodes": a.nodes, "group": a.group,
                      "solves_per_s": a.batch / best, "seconds": best,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}))


if __name__ == "__main__":
    main()
