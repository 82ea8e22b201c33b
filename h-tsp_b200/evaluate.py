"""The reference's evaluation loop (evaluate.py:74-101: VecEnv.reset, then `while not done: action = policy(state);
state, _, done = env.step(action, rl_solver)`) on the device-resident pipeline: B200VecEnv (k-NN reset, fragment
selection, scoring), B200UpperPolicy or the reference's random stand-in, B200RLSolver (the path-solver hot path).
Used by tools/solve_tsp.py, bench.py (BASELINE configs 3 / 4 as secondary keys) and the end-to-end parity test."""
from __future__ import annotations

import time
from typing import Dict, Optional

import torch

from .vec_env import B200VecEnv

# evaluate.py:25-41 (k, frag_len, max_new_nodes, max_improvement_step, x8Aug via B200RLSolver, 200 POMO starts)
EVAL_SETTINGS = dict(k=40, frag_len=200, max_new_nodes=190, max_improvement_step=0)


@torch.no_grad()
def construct_tours(x: torch.Tensor, hook, policy=None, settings: Optional[Dict] = None, shard_solver: bool = False,
                    check: bool = True, seed: Optional[int] = None) -> Dict:
    """x [E,Ntot,2] on the device -> dict(tours list[list[int]], lengths [E] cpu f32, seconds, env_steps, breakdown_ms).
    policy: callable(state [E,Ntot,8]) -> actions [E,2] in [-1,1]^2, or None for VecEnv.random_action (seeded by
    `seed` through torch.manual_seed, as evaluate.py does)."""
    cfg = dict(EVAL_SETTINGS if settings is None else settings)
    dev = x.device
    env = B200VecEnv(**cfg)
    if seed is not None:
        torch.manual_seed(seed)
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    state = env.reset(x)
    torch.cuda.synchronize(dev)
    t_reset = time.perf_counter() - t0
    steps = 0
    t_pol = t_frag = t_solve = t_score = 0.0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    while not env.done:
        ev[4].record()
        act = env.random_action() if policy is None else policy(state)
        ev[0].record()
        active = ~env.dones
        idx = torch.nonzero(active).squeeze(1)
        frag, _, _, _ = env.fragments(act, active)
        ev[1].record()
        if shard_solver:
            from . import sharding
            paths, _ = sharding.solve_sharded(hook, env.x.index_select(0, idx), frag.index_select(0, idx))
            status = torch.zeros(1, dtype=torch.int32, device=dev)
        else:
            paths, _, status, _ = hook.solve_device(env.x.index_select(0, idx), frag.index_select(0, idx), None)
        ev[2].record()
        s = env.state
        s._paths.zero_()
        s._path_len.zero_()
        s._paths.index_copy_(0, idx, paths)
        s._path_len.index_fill_(0, idx, env.frag_len)
        state, _ = s.step_device(check=check)
        ev[3].record()
        torch.cuda.synchronize(dev)
        if check:
            assert bool((status == 0).all()), "invalid path(s) returned by the rollout"
        t_pol += ev[4].elapsed_time(ev[0])
        t_frag += ev[0].elapsed_time(ev[1])
        t_solve += ev[1].elapsed_time(ev[2])
        t_score += ev[2].elapsed_time(ev[3])
        steps += 1
    total = time.perf_counter() - t0
    E, N = env.E, env.N
    tours = [env.state.current_tour(e) for e in range(E)]
    for t in tours:
        assert len(t) == N and len(set(t)) == N, "constructed tour is not a permutation of the cities"
    return {"tours": tours, "lengths": env.state.cur_len.cpu(), "seconds": total, "env_steps": steps,
            "breakdown_ms": {"knn_reset": t_reset * 1e3, "upper_policy": t_pol, "fragment_selection": t_frag,
                             "path_solver_hook": t_solve, "scoring": t_score}}
