"""TEST INFRASTRUCTURE ONLY -- the UNMODIFIED reference pipeline end to end (evaluate.py:74-101): h_tsp.VecEnv +
h_tsp.RLSolver over the reference PathSolver + the upper-level policy (rl_models.IMPALAEncoder / ActorPPO as
HTSP_PPO.forward composes them, h_tsp.py:1272-1276), on CPU, with the evaluate.py settings (k = 40, fragments of
200, <= 190 new cities, x8Aug, 200 POMO starts, no improvement steps) and SEEDED RANDOM-INIT weights
(synthetic_state_dict(1234, 6), synthetic_policy_state_dict(1234)) on TSP-1000 instances.  Records the instances and
the final tour lengths into tests/golden/e2e_pipeline_golden.npz; tests/test_e2e_pipeline.py runs the device-resident
pipeline on the same instances and weights and compares the tour lengths statistically (the POMO starts come from
torch.randperm on different generators, so individual tours differ).

    python oracle/make_golden_e2e.py          # ~2 minutes on 8 cores

Two runs with different torch seeds are stored: their spread is the run-to-run noise of the reference itself and sets
the tolerance of the test."""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
import htsp_b200  # noqa: E402,F401
from htsp_b200.weights import synthetic_state_dict  # noqa: E402

CITIES, ENVS, G = 1000, 4, 200
SETTINGS = dict(k=40, frag_len=200, max_new_nodes=190, max_improvement_step=0)


def main():
    H = ref_loader.load_h_tsp_module()
    M = ref_loader.load_rl_models_module()
    import rl_utils
    pol = importlib.import_module("htsp_b200.upper_policy")
    model = ref_loader.build_reference_path_solver(synthetic_state_dict(1234, 6), n_layers=6)
    usd = pol.synthetic_policy_state_dict(1234)
    enc = M.IMPALAEncoder(input_dim=8, embedding_dim=128)                                  # h_tsp.py:949-952
    act = M.ActorPPO(state_dim=128, mid_dim=128, action_dim=2, init_a_std_log=-1)          # h_tsp.py:959-964
    enc.load_state_dict({k[len("encoder."):]: v for k, v in usd.items() if k.startswith("encoder.")}, strict=True)
    act.load_state_dict({k[len("actor."):]: v for k, v in usd.items() if k.startswith("actor.")}, strict=False)
    enc.eval()
    act.eval()
    x = torch.rand(ENVS, CITIES, 2, generator=torch.Generator().manual_seed(20261020))
    out = {"x": x.numpy(), "settings": np.array([SETTINGS["k"], SETTINGS["frag_len"], SETTINGS["max_new_nodes"], G],
                                                dtype=np.int32)}
    for run, seed in enumerate((11, 22)):
        torch.manual_seed(seed)
        np.random.seed(seed)
        solver = H.RLSolver(model, sample_size=G)
        venv = H.VecEnv(**SETTINGS)
        fb = rl_utils.FragmengBuffer(10000, SETTINGS["frag_len"], 2)
        t0 = time.time()
        with torch.no_grad():
            s = venv.reset(x)
            steps = 0
            while not venv.done:
                a = act(enc(s)).detach()                                                   # HTSP_PPO.forward
                s, r, d, info = venv.step(a, solver, frag_buffer=fb)
                steps += 1
        lens = np.array([e.state.current_tour_len.item() for e in venv.envs], dtype=np.float64)
        for e in venv.envs:
            t = list(e.state.current_tour)
            assert sorted(t) == list(range(CITIES))
        out[f"lengths_run{run}"] = lens
        out[f"steps_run{run}"] = np.int32(steps)
        print(f"run {run} (seed {seed}): {steps} env steps in {time.time() - t0:.1f} s, tour lengths {lens.round(3).tolist()}, "
              f"mean {lens.mean():.4f}")
    dst = os.path.join(ROOT, "tests", "golden", "e2e_pipeline_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
