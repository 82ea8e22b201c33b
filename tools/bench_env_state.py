#!/usr/bin/env python
"""Measurement of K6 (device-resident environment state): one
`Env._step` for E environments of Ntot cities with 200-city paths, inputs resident in HBM.
Prints one JSON line: env steps/s, the HBM roofline fraction of the two kernels against their
algorithmic bytes, and the oracle port (numpy restatement of the reference arithmetic) on the host.

  python tools/bench_env_state.py [--envs 16] [--cities 10000] [--steps 30]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import htsp_b200  # noqa: E402,F401
from htsp_b200.env_state import B200VecState  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16)
    ap.add_argument("--cities", type=int, default=10000)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--k", type=int, default=40)
    a = ap.parse_args()
    E, N, k, P = a.envs, a.cities, a.k, 200
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(E, N, 2, generator=g).cuda()
    knn = torch.empty(E, N, k, dtype=torch.int64, device="cuda")
    for e in range(E):
        d = torch.cdist(x[e].double(), x[e].double())
        d.fill_diagonal_(float("inf"))
        knn[e] = d.topk(k, largest=False).indices
        del d
    v = B200VecState(x, knn, [[0, int(knn[e, 0, 0])] for e in range(E)], max_path_len=P)
    rng = np.random.default_rng(0)

    def make_paths():
        paths = []
        for e in range(E):
            tour = v.current_tour(e)
            unseen = np.setdiff1d(np.arange(N), tour)
            i = int(rng.integers(0, len(tour)))
            w = min(len(tour), 10)
            window = [tour[(i + j) % len(tour)] for j in range(w)]
            new = [int(c) for c in rng.permutation(unseen)[:P - w]]
            paths.append([window[0]] + [int(c) for c in rng.permutation(window[1:-1] + new)] + [window[-1]])
        return paths

    for _ in range(3):                                   # warm-up (also grows the tours)
        v.step(make_paths())
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms, bytes_alg = [], []
    for _ in range(a.steps):
        v.set_paths(make_paths())                        # host -> device path ids, outside the timed region
        T = v.tour_len.cpu().numpy().astype(np.int64)
        torch.cuda.synchronize()
        ev0.record()
        v.step_device(check=False)
        ev1.record()
        torch.cuda.synchronize()
        ms.append(ev0.elapsed_time(ev1))
        T2 = v.tour_len.cpu().numpy().astype(np.int64)
        A = v.available.sum(dim=1).cpu().numpy().astype(np.int64)
        # splice: copy T ids (read+write) + index rebuild (T + T2 ints); update: T2*(8 id + 8 xy + 16 row)
        # + length T2*(8+8) + masks A*k*(8+1) + state N*(8+1+1+16+32)
        b = (T + T2) * 8 + (T + T2) * 4 + T2 * 32 + T2 * 16 + A * k * 9 + N * 58
        bytes_alg.append(int(b.sum()))
    t = float(np.median(ms)) * 1e-3
    gbs = float(np.median(bytes_alg)) / t / 1e9
    peak = 6555.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    # oracle port on the host (numpy restatement of the reference arithmetic, one environment)
    from oracle import env_state_oracle as eo
    xe, ke = x[0].cpu().numpy(), knn[0].cpu().numpy()
    st = eo.LargeStateOracle(xe, ke, [0, int(ke[0, 0])])
    t0 = time.perf_counter()
    n_cpu = 0
    while time.perf_counter() - t0 < 3.0 and st.current_num_cities < N:
        tour = st.current_tour
        unseen = np.setdiff1d(np.arange(N), tour)
        w = min(len(tour), 10)
        window = tour[:w]
        p = [window[0]] + window[1:-1] + [int(c) for c in unseen[:P - w]] + [window[-1]]
        eo.env_step(st, p)
        n_cpu += 1
    cpu_rate = n_cpu / (time.perf_counter() - t0)
    print(json.dumps({
        "metric": "large-TSP environment steps/s (Env._step bookkeeping, 200-city paths)", "unit": "env steps/s",
        "value": E / t, "ms_per_batched_step": t * 1e3,
        "config": {"workload": f"{E} environments x TSP-{N}, k={k}, paths of {P} cities, two launches per step"},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                     "traffic": None, "note": "E CTAs of 1024 threads: a latency-bound launch, not a bandwidth-bound one"},
        "cpu_baseline": {"value": cpu_rate, "unit": "env steps/s", "cores": 1, "kind": "port",
                         "sample": f"{n_cpu} steps of one TSP-{N} environment, numpy restatement (O(T) per step; "
                                   "the reference itself gathers an Ntot x Ntot matrix, 57 ms per call at 10000)"}}))


if __name__ == "__main__":
    main()
