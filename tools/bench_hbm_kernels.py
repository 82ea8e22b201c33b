#!/usr/bin/env python
"""HBM-roofline measurement of the byte/integer kernels around the rollout (K1 extract/normalise/
augment, K4 best-of selection + path orientation, K5 tour length): CUDA-event timing with inputs
resident in HBM and larger than L2 (or L2 flushed), achieved = ALGORITHMIC bytes / time against the
measured copy bandwidth of MEASURED_PEAKS.json.  One JSON line per kernel.

  python tools/bench_hbm_kernels.py [--batch 65536]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import htsp_b200  # noqa: E402,F401
from htsp_b200 import _abi  # noqa: E402


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def timed(fn, reps=10, flush=None):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        fn()
    ms = []
    for _ in range(reps):
        if flush is not None:
            flush.add_(1)                     # 512 MB write: evicts the 126 MB L2
        torch.cuda.synchronize()
        ev0.record()
        fn()
        ev1.record()
        torch.cuda.synchronize()
        ms.append(ev0.elapsed_time(ev1))
    ms.sort()
    return ms[len(ms) // 2] * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    a = ap.parse_args()
    lib = _abi.lib()
    peak, src = peak_hbm()
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.zeros(128 * 1024 * 1024, dtype=torch.int32, device="cuda")
    out = []
    B, N, G = a.batch, 200, 200

    # ---- K1: gather + min/max normalise + 8-fold augmentation (h_tsp.py:176-191, utils.py:37-56)
    Ntot, E = 10000, 64                       # 64 TSP-10000 instances, B fragments drawn over them
    data = torch.rand(E, Ntot, 2, device="cuda")
    datab = data.repeat(B // E, 1, 1) if B // E * E * Ntot * 8 < 8e9 else None
    Bk = B if datab is not None else E
    datab = datab if datab is not None else data
    frag = torch.stack([torch.randperm(Ntot, device="cuda")[:N] for _ in range(64)]).repeat(Bk // 64, 1).contiguous()
    x_new = torch.empty(Bk, N, 2, device="cuda")
    scale = torch.empty(Bk, device="cuda")
    x_aug = torch.empty(8 * Bk, N, 2, device="cuda")
    for n_aug, aug in ((1, None), (8, x_aug)):
        def k1():
            _abi.check(lib.htsp_extract_normalize(datab.data_ptr(), frag.data_ptr(), Bk, Ntot, N, n_aug,
                                                  x_new.data_ptr(), scale.data_ptr(), _abi.ptr(aug), st), "k1")
        t = timed(k1, flush=flush)
        byt = Bk * (N * (8 + 8 + 8) + 4) + (Bk * N * 8 * 8 if n_aug == 8 else 0)
        out.append({"kernel": f"K1 extract_normalize (n_aug={n_aug})", "subproblems": Bk, "seconds": t,
                    "algorithmic_bytes": byt,
                    "roofline": {"bound": "hbm", "achieved": byt / t / 1e9, "peak": peak, "unit": "GB/s",
                                 "frac": byt / t / 1e9 / peak, "peak_source": src},
                    "note": "gathered coordinates are 8-byte random reads (32-byte sectors): DRAM moves 4x the "
                            "algorithmic gather bytes"})

    # ---- K4: best-of-G selection + orientation (train_path_solver.py:290-306, h_tsp.py:206-226)
    Bs = min(B, 16384)
    reward = -torch.rand(Bs, G, device="cuda")
    pi = torch.stack([torch.randperm(N, device="cuda") for _ in range(64)]).int()
    pi = pi[torch.randint(0, 64, (Bs * G,), device="cuda")].reshape(Bs, G, N).contiguous()
    best_len = torch.empty(Bs, device="cuda")
    best_row = torch.empty(Bs, dtype=torch.int32, device="cuda")
    best_pi = torch.empty(Bs, N, dtype=torch.int64, device="cuda")

    def k4():
        _abi.check(lib.htsp_select_best(reward.data_ptr(), pi.data_ptr(), 1, Bs, G, N, best_len.data_ptr(),
                                        best_row.data_ptr(), best_pi.data_ptr(), st), "k4")
    t = timed(k4, flush=flush)
    byt = Bs * (G * 4 + N * 4 + N * 8 + 8)
    out.append({"kernel": "K4 select_best (noAug_nTraj)", "subproblems": Bs, "seconds": t, "algorithmic_bytes": byt,
                "roofline": {"bound": "hbm", "achieved": byt / t / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": byt / t / 1e9 / peak, "peak_source": src},
                "note": "reads G rewards and ONE winning tour per subproblem out of a G x N x 4 B tour block: "
                        "small, latency-bound"})

    # ---- K5: tour length from coordinates (rl_utils.py:87-102)
    Et, T = 256, 10000
    coords = torch.rand(Et, T, 2, device="cuda")
    tours = torch.stack([torch.randperm(T, device="cuda") for _ in range(Et)])
    tl = torch.full((Et,), T, dtype=torch.int32, device="cuda")
    res = torch.empty(Et, device="cuda")

    def k5():
        _abi.check(lib.htsp_tour_length(coords.data_ptr(), tours.data_ptr(), tl.data_ptr(), Et, T, T,
                                        res.data_ptr(), st), "k5")
    t = timed(k5, flush=flush)
    byt = Et * T * 16
    out.append({"kernel": "K5 tour_length (TSP-10000 tours)", "tours": Et, "seconds": t, "algorithmic_bytes": byt,
                "roofline": {"bound": "hbm", "achieved": byt / t / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": byt / t / 1e9 / peak, "peak_source": src},
                "note": "one CTA of 1024 threads per tour, four edges in flight per thread, the edge end from the next lane by shuffle; random 8-byte coordinate gathers"})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
