"""TEST INFRASTRUCTURE ONLY -- records the environment scoring step of the UNMODIFIED reference
(h_tsp.py VecEnv / Env._step / LargeState.move_to) on real episodes into
tests/golden/env_state_golden.npz.  Run in the build container: python oracle/make_golden_env.py

The episodes are driven by random upper-level actions and a nearest-neighbour path solver wrapped
in a subclass of the reference's RLSolver (VecEnv.step only accepts RLSolver / LKHSolver paths,
h_tsp.py:764-782); what is recorded is independent of the solver: every new_path handed to
Env._step and the complete state after it."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def main():
    H = ref_loader.load_h_tsp_module()
    torch.manual_seed(7)
    np.random.seed(7)

    class NNPathSolver(H.RLSolver):
        def __init__(self):
            pass

        def solve(self, data, fragment, frag_buffer=None):
            paths = []
            for b in range(fragment.shape[0]):
                p, _ = H.GreedySolver().solve(data[b].cpu().numpy(), fragment[b].cpu().tolist())
                paths.append([int(c) for c in p])
            return paths, None

    out = {}
    for tag, (N, frag_len, max_new, k) in {"n300": (300, 40, 30, 8), "n1000": (1000, 200, 190, 40)}.items():
        x = torch.rand(2, N, 2)
        venv = H.VecEnv(k=k, frag_len=frag_len, max_new_nodes=max_new, max_improvement_step=2)
        venv.reset(x)
        env = venv.envs[0]
        st = env.state
        out[f"{tag}_x"] = x[0].numpy()
        out[f"{tag}_knn"] = st.knn_neighbor.numpy()
        out[f"{tag}_init_tour"] = np.array(st.current_tour, dtype=np.int64)
        out[f"{tag}_init_len"] = np.float32(st.current_tour_len)
        out[f"{tag}_init_state"] = st.to_tensor().numpy()
        paths, tours, tour_lens, lens, rewards, states, wrap = [], [], [], [], [], [], []
        acts, frags, newc, starts = [], [], [], []
        solver = NNPathSolver()
        orig_step = env._step

        def rec_step(p, *a, _orig=orig_step, **kw):
            old = list(env.state.current_tour)
            s_i, e_i = (old.index(p[0]), old.index(p[-1])) if old else (0, 1)
            res = _orig(p, *a, **kw)
            paths.append(np.array(p, dtype=np.int64))
            t = np.full(N, -1, dtype=np.int64)
            t[:len(env.state.current_tour)] = env.state.current_tour
            tours.append(t)
            tour_lens.append(len(env.state.current_tour))
            lens.append(np.float32(env.state.current_tour_len))
            rewards.append(np.float32(res[1]))
            states.append(res[0].numpy().copy())
            wrap.append(e_i < s_i)
            return res

        env._step = rec_step
        orig_frag = env.get_fragment_knn

        def rec_frag(a, _orig=orig_frag):
            res = _orig(a)
            acts.append(a.numpy().astype(np.float32).copy())          # the SCALED action in [0,1]^2
            frags.append(res[0].numpy().astype(np.int64))
            nc = np.full(frag_len, -1, dtype=np.int64)
            nc[:len(res[1])] = res[1].numpy()
            newc.append(nc)
            starts.append(res[2].numpy().astype(np.float32))
            return res

        env.get_fragment_knn = rec_frag
        while not venv.done:
            venv.step(venv.random_action(), solver)
        L = max(len(p) for p in paths)
        pp = np.full((len(paths), L), -1, dtype=np.int64)
        for i, p in enumerate(paths):
            pp[i, :len(p)] = p
        out[f"{tag}_actions"] = np.stack(acts)
        out[f"{tag}_fragments"] = np.stack(frags)
        out[f"{tag}_new_cities"] = np.stack(newc)
        out[f"{tag}_start_coord"] = np.stack(starts)
        out[f"{tag}_cfg"] = np.array([frag_len, max_new, k], dtype=np.int32)
        out[f"{tag}_paths"] = pp
        out[f"{tag}_path_len"] = np.array([len(p) for p in paths], dtype=np.int32)
        out[f"{tag}_tours"] = np.stack(tours)
        out[f"{tag}_tour_len"] = np.array(tour_lens, dtype=np.int32)
        out[f"{tag}_len"] = np.array(lens, dtype=np.float32)
        out[f"{tag}_reward"] = np.array(rewards, dtype=np.float32)
        # the full [N,8] state after a few of the steps (all steps for the small instance)
        keep = list(range(len(states))) if N <= 300 else [0, len(states) // 2, len(states) - 1]
        out[f"{tag}_state_steps"] = np.array(keep, dtype=np.int32)
        out[f"{tag}_states"] = np.stack([states[i] for i in keep])
        print(f"{tag}: {len(paths)} env steps, {sum(wrap)} wrap-around splices, final tour {tour_lens[-1]} cities, "
              f"length {float(lens[-1]):.4f}")
        env.debug_check_tour_len()
    dst = os.path.join(ROOT, "tests", "golden", "env_state_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
