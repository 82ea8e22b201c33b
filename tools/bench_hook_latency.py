#!/usr/bin/env python
"""Latency of one lower-level hook call (B200RLSolver.solve_device: gather/normalise, x8 augmentation, encoder,
200-start greedy rollout, best-of selection, path orientation) for E subproblems = 8E rollout instances, i.e. one
environment step of E large-TSP environments (evaluate.py: E = 16).  With fewer instances than SMs the rollout
kernel splits the POMO rows of an instance over several CTAs (csrc/decode_tc.cu, tcd_choose_split);
HTSP_DECODE_SPLIT forces the number of parts.  One JSON line per E.

  python tools/bench_hook_latency.py [--envs 1,2,4,8,16] [--reps 5]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import htsp_b200  # noqa: E402,F401
from htsp_b200 import B200PathSolver, B200RLSolver  # noqa: E402
from htsp_b200.weights import synthetic_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default="1,2,4,8,16")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--mode", default="x3")
    a = ap.parse_args()
    dev = torch.device("cuda")
    model = B200PathSolver(state_dict=synthetic_state_dict(1234, 6), mode=a.mode, device=dev, graph_size=200)
    model.first_action_override = torch.randperm(200, generator=torch.Generator().manual_seed(7))[:200]
    hook = B200RLSolver(model, sample_size=200)
    g = torch.Generator().manual_seed(3)
    for E in [int(e) for e in a.envs.split(",")]:
        data = torch.rand(E, 10000, 2, generator=g).to(dev)
        frag = torch.stack([torch.randperm(10000, generator=g)[:200] for _ in range(E)]).to(dev)
        res, ref = {}, None
        for split in ("1", "2", "4", "8", None):
            if split is None:
                os.environ.pop("HTSP_DECODE_SPLIT", None)
            else:
                os.environ["HTSP_DECODE_SPLIT"] = split
            paths, lens, status, _ = hook.solve_device(data, frag, None)
            torch.cuda.synchronize()
            if ref is None:
                ref = (paths.clone(), lens.clone())
            same = bool(torch.equal(paths, ref[0])) and bool(torch.equal(lens, ref[1]))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.reps):
                hook.solve_device(data, frag, None)
            e1.record()
            torch.cuda.synchronize()
            res["auto" if split is None else f"split{split}"] = {"ms": e0.elapsed_time(e1) / a.reps,
                                                                 "identical_to_unsplit": same}
        print(json.dumps({"metric": "lower-level hook call latency (x8Aug, 200 POMO starts, 200-node subproblems)",
                          "unit": "ms", "subproblems": E, "rollout_instances": 8 * E, "arith": a.mode, **res}), flush=True)


if __name__ == "__main__":
    main()
