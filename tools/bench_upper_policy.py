#!/usr/bin/env python
"""Measurement of K8 (upper-level policy forward): HTSP_PPO.forward for E
environment states of Ntot cities resident in HBM -> E actions.  Prints one JSON line per Ntot: forward
calls/s, ms per call, the share of the 3x3 convolutions, the fp32 FMA rate reached against the nominal
CUDA-core peak (148 SMs x 128 lanes x 2 flop x SM clock; MEASURED_PEAKS.json has no fp32 entry), and the
oracle port (torch CPU fp32 + numpy pillar stage) on the host cores beside it.

  python tools/bench_upper_policy.py [--envs 16] [--cities 1000,10000] [--steps 50]"""
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
from htsp_b200 import _abi  # noqa: E402
from htsp_b200.upper_policy import B200UpperPolicy, synthetic_policy_state_dict, DEPTHS, FLAT_DIM, GRID  # noqa: E402


def algorithmic_flops(embed=128, mid=128, adim=2):
    """FLOP = 2 MAC of the dense part of one forward (convs, fc, actor); the pillar stage adds 2*12*C0 per city."""
    c0, f, hw, cin = embed // 2, 0, GRID, embed // 2
    for d in DEPTHS:
        f += 2 * hw * hw * 9 * cin * d
        hw = (hw - 1) // 2 + 1
        f += 4 * 2 * hw * hw * 9 * d * d
        cin = d
    f += 2 * FLAT_DIM * embed + 2 * (embed * mid + 2 * mid * mid + mid * adim)
    return f, 2 * 12 * c0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16)
    ap.add_argument("--cities", default="1000,10000")
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    sd = synthetic_policy_state_dict(1234)
    pol = B200UpperPolicy(sd, "cuda:0")
    lib = _abi.lib()
    rs = np.random.RandomState(0)
    sm_mhz = 1965          # max SM clock of this pool's B200s (bench.py's clocks line)
    for N in [int(c) for c in a.cities.split(",")]:
        E = a.envs
        s = rs.rand(E, N, 8).astype(np.float32)
        s[:, :, 2] = s[:, :, 2] < 0.6
        s[:, :, 3] = (s[:, :, 3] < 0.4) * s[:, :, 2]
        sdev = torch.from_numpy(s).cuda()
        for _ in range(5):
            pol(sdev)
        torch.cuda.synchronize()
        l0 = lib.htsp_launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(a.steps):
            act = pol(sdev)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / a.steps
        launches = (lib.htsp_launch_count() - l0) // a.steps
        dense, per_city = algorithmic_flops()
        flops = E * (dense + per_city * N)
        peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        # host baseline: the oracle port on all host cores (bounded: 3 calls)
        sys.path.insert(0, ROOT)
        from oracle import upper_policy_oracle as uo
        uo.policy_forward(s[:2], sd)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            _, _, o_act = uo.policy_forward(s, sd)
        cpu_ms = (time.perf_counter() - t0) / reps * 1e3
        err = float(np.abs(act.cpu().numpy() - o_act).max())
        print(json.dumps({
            "metric": "upper-level policy forward (IMPALAEncoder + ActorPPO), environment states/s", "unit": "states/s",
            "value": E / (ms * 1e-3), "ms_per_call": ms, "envs": E, "cities": N, "gpu_launches_per_call": int(launches),
            "dtype": "f32", "higher_is_better": True,
            "roofline": {"bound": "fp32 FMA (CUDA cores)", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak,
                         "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / peak,
                         "peak_source": f"nominal 148 SM x 128 FMA x 2 x {sm_mhz} MHz",
                         "algorithmic_flops_per_call": flops},
            "cpu_baseline": {"value": E / (cpu_ms * 1e-3), "unit": "states/s", "ms_per_call": cpu_ms, "kind": "port",
                             "cores": torch.get_num_threads(),
                             "sample": f"{reps} calls of the oracle port (numpy pillar stage + torch CPU fp32 CNN)"},
            "max_abs_action_diff_vs_oracle": err}))


if __name__ == "__main__":
    main()
