#!/usr/bin/env python
"""End-to-end large-TSP construction on the device-resident pipeline (BASELINE configs 3 / 4 shape):
B200VecEnv (k-NN reset, fragment selection, scoring) + B200RLSolver (the path-solver hot path), with
the reference's evaluate.py settings (k=40, frag_len=200, max_new_nodes=190, max_improvement_step=0,
x8Aug, 200 POMO starts, bsz 16).  --policy upper runs the upper-level policy forward on the device too
(B200UpperPolicy, evaluate.py:84); --policy random uses the reference's random stand-in
(VecEnv.random_action, evaluate.py:86).  Trained checkpoints are not available offline: upper and
lower weights are random init unless --ckpt / --upper-ckpt give reference state_dicts -- so the printed
GAP IS NOT the paper's number; the solve TIME and the per-step breakdown are what this tool measures.

  python tools/solve_tsp.py [--cities 10000] [--envs 16] [--mode x3] [--policy upper|random]
                            [--ckpt path_solver.ckpt] [--upper-ckpt htsp_ppo.ckpt]

Under torchrun (one rank per GPU, BASELINE config 4) every rank constructs its own --envs instances
(instances are independent: no data-path collective); the ranks meet at a barrier before and after, the
reported time is the maximum over ranks and the tour lengths are all-gathered for the mean:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/solve_tsp.py"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import htsp_b200  # noqa: E402,F401
from htsp_b200 import B200PathSolver, B200RLSolver, sharding  # noqa: E402
from htsp_b200.evaluate import construct_tours  # noqa: E402
from htsp_b200.weights import synthetic_state_dict  # noqa: E402
from htsp_b200.upper_policy import B200UpperPolicy, synthetic_policy_state_dict  # noqa: E402

OPTIMA = {1000: 23.1182, 10000: 71.7778}      # rl4cop/eval.py:116 (mean optimal length, uniform instances)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cities", type=int, default=10000)
    ap.add_argument("--envs", type=int, default=16)
    ap.add_argument("--mode", default="x3")
    ap.add_argument("--group", type=int, default=200)
    ap.add_argument("--ckpt", default=None)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--policy", choices=("upper", "random"), default="upper")
    ap.add_argument("--upper-ckpt", default=None)
    ap.add_argument("--shard-solver", action="store_true",
                    help="under torchrun: all ranks construct the SAME --envs instances, the subproblems of every "
                         "environment step are sharded over the ranks (latency of one batch instead of throughput)")
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if a.ckpt:
        sd = torch.load(a.ckpt, map_location="cpu")
        sd = sd.get("state_dict", sd)
    else:
        sd = synthetic_state_dict(1234, 6)
    model = B200PathSolver(state_dict=sd, mode=a.mode, device=dev, graph_size=200)
    hook = B200RLSolver(model, sample_size=a.group)
    policy = None
    if a.policy == "upper":
        if a.upper_ckpt:
            usd = torch.load(a.upper_ckpt, map_location="cpu")
            usd = usd.get("state_dict", usd)
        else:
            usd = synthetic_policy_state_dict(a.seed)
        policy = B200UpperPolicy(usd, dev, embedding_dim=usd["encoder.cnn.fc.weight"].shape[0],
                                 mid_dim=usd["actor.net.0.0.weight"].shape[0])
    shard = a.shard_solver and world > 1
    g = torch.Generator().manual_seed(a.seed + (0 if shard else 7919 * rank))   # own instances per rank unless sharded
    x = torch.rand(a.envs, a.cities, 2, generator=g).to(dev)
    torch.manual_seed(a.seed)
    # warm-up on a small instance (library load, workspace allocation, first-launch overheads)
    construct_tours(torch.rand(2, 600, 2, generator=g).to(dev), hook, policy)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    res = construct_tours(x, hook, policy, shard_solver=shard)
    total, steps = res["seconds"], res["env_steps"]
    bd = res["breakdown_ms"]
    t_reset, t_pol, t_frag, t_solve, t_score = (bd["knn_reset"] / 1e3, bd["upper_policy"], bd["fragment_selection"],
                                                bd["path_solver_hook"], bd["scoring"])
    lens = res["lengths"].to(dev)
    if world > 1:
        tt = torch.tensor([total], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total = float(tt)
        if not shard:
            all_lens = torch.empty(world * a.envs, device=dev, dtype=lens.dtype)
            dist.all_gather_into_tensor(all_lens, lens.contiguous())
            lens = all_lens
    lens = lens.cpu()
    opt = OPTIMA.get(a.cities)
    n_total = a.envs if shard else a.envs * world
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    print(json.dumps({
        "metric": f"TSP-{a.cities} end-to-end construction time (device-resident H-TSP loop)", "unit": "s",
        "value": total, "higher_is_better": False, "n_gpus": world,
        "multi_gpu": None if world == 1 else ("subproblems of every step sharded over the ranks (same instances)"
                                              if shard else "independent instances per rank"), "n_instances": n_total, "env_steps": steps,
        "seconds_per_instance": total / n_total,
        "breakdown_ms": {"knn_reset": t_reset * 1e3, "upper_policy": t_pol, "fragment_selection": t_frag, "path_solver_hook": t_solve,
                         "scoring": t_score},
        "mean_tour_length": float(lens.mean()),
        "gap_vs_rl4cop_eval_optimum": None if opt is None else float(lens.mean() / opt - 1),
        "upper_policy": a.policy if a.policy == "random" else ("checkpoint" if a.upper_ckpt else "random-init weights"),
        "caveat": ("random upper-level actions" if policy is None else
                   ("checkpoint" if a.upper_ckpt else "random-init") + " upper-level policy") + " and " +
                  ("checkpoint" if a.ckpt else "random-init") +
                  " lower-level weights: the gap is not the paper's; time and breakdown are the measurement",
        "config": {"k": 40, "frag_len": 200, "max_new_nodes": 190, "val_type": "x8Aug_2Traj", "pomo_starts": a.group,
                   "arith": a.mode}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
