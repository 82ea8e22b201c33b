"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, via oracle/ref_loader.py) on seeded inputs.  Run in the build container:
    python oracle/make_golden.py
The reference ships no golden vectors for this path, so these files are
the pin: tests compare both the oracle restatement and the CUDA path against them."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import htsp_b200  # noqa: E402,F401
from htsp_b200.weights import synthetic_state_dict  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
LOGIT_STEPS = [0, 1, 2, 50, 100, 150, 196, 197]


def forward_case(model, seed, B, N, G, val_type):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, N, 2, generator=g)
    src = torch.zeros(B, 1, dtype=torch.long)
    tgt = torch.full((B, 1), N - 1, dtype=torch.long)
    torch.manual_seed(seed)
    first = torch.randperm(N)[:G].clone()
    torch.manual_seed(seed)
    with torch.no_grad():
        length, pi = model(x, src, tgt, val_type=val_type, return_pi=True, group_size=G)
    return dict(x=x.numpy(), first=first.numpy().astype(np.int32), length=length.numpy(),
                pi=pi.numpy().astype(np.int32))


def main():
    os.makedirs(OUT, exist_ok=True)
    tps = ref_loader.load_path_solver_module()
    h_tsp = ref_loader.load_h_tsp_module()
    import rl_utils  # reference module (sys.path set by the loader)

    out = {}
    sd6 = synthetic_state_dict(1234, 6)
    m6 = ref_loader.build_reference_path_solver(sd6, n_layers=6)

    # --- PathSolver.forward, all three val_type branches, N=50 ---------------------------
    for tag, vt in (("aug", "x8Aug_2Traj"), ("ntraj", "noAug_nTraj"), ("1traj", "noAug_1Traj")):
        c = forward_case(m6, 11, 3, 50, 6, vt)
        for k, v in c.items():
            out[f"fwd_{tag}_{k}"] = v

    # --- full-size rollout with per-step logits, N=200, G=8, no augmentation --------------
    g = torch.Generator().manual_seed(21)
    x = torch.rand(2, 200, 2, generator=g) * 0.9 + 0.05
    src = torch.zeros(2, 1, dtype=torch.long)
    tgt = torch.full((2, 1), 199, dtype=torch.long)
    torch.manual_seed(21)
    first = torch.randperm(200)[:8].clone()
    torch.manual_seed(21)
    with torch.no_grad(), ref_loader.SoftmaxTap() as tap:
        length, pi = m6(x, src, tgt, val_type="noAug_1Traj", return_pi=True, group_size=8)
        emb = m6.encoder(x)
    assert len(tap.logits) == 198
    out["n200_x"] = x.numpy()
    out["n200_first"] = first.numpy().astype(np.int32)
    out["n200_pi"] = pi.numpy().astype(np.int32)
    out["n200_length"] = length.numpy()
    out["n200_logit_steps"] = np.array(LOGIT_STEPS, dtype=np.int32)
    out["n200_logits"] = torch.stack([tap.logits[s] for s in LOGIT_STEPS]).numpy()
    out["n200_emb_head"] = emb[:, :8, :].numpy()

    # --- a different depth (3 layers) with G = N and src/tgt not at the ends --------------
    sd3 = synthetic_state_dict(77, 3)
    m3 = ref_loader.build_reference_path_solver(sd3, n_layers=3, graph_size=20, group_size=20)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 20, 2, generator=g)
    src = torch.tensor([[3], [7]])
    tgt = torch.tensor([[11], [0]])
    torch.manual_seed(5)
    first = torch.randperm(20)[:20].clone()
    torch.manual_seed(5)
    with torch.no_grad():
        length, pi = m3(x, src, tgt, val_type="noAug_nTraj", return_pi=True, group_size=20)
    out["l3_x"], out["l3_src"], out["l3_tgt"] = x.numpy(), src.numpy().astype(np.int32), tgt.numpy().astype(np.int32)
    out["l3_first"] = first.numpy().astype(np.int32)
    out["l3_length"], out["l3_pi"] = length.numpy(), pi.numpy().astype(np.int32)

    # --- the RLSolver hook (h_tsp.py:165-228) --------------------------------------------
    g = torch.Generator().manual_seed(31)
    data = torch.rand(3, 400, 2, generator=g)
    frag = torch.stack([torch.randperm(400, generator=g)[:50] for _ in range(3)])
    fb = rl_utils.FragmengBuffer(8, 50, 2)
    solver = h_tsp.RLSolver(m6, 6)
    torch.manual_seed(31)
    first = torch.randperm(50)[:6].clone()
    torch.manual_seed(31)
    paths, lens = solver.solve(data, frag, fb)
    out["hook_data"], out["hook_fragment"] = data.numpy(), frag.numpy()
    out["hook_first"] = first.numpy().astype(np.int32)
    out["hook_paths"] = np.array(paths, dtype=np.int64)
    out["hook_lengths"] = np.asarray(lens, dtype=np.float32)
    out["hook_x_new"] = fb.frag_buffer[:3].numpy()

    # --- scoring: get_tour_distance over the fp64->fp32 dist matrix (rl_utils.py:87-102) ---
    g = torch.Generator().manual_seed(41)
    xs = torch.rand(1000, 2, generator=g)
    dm = torch.cdist(xs.double() * 1000, xs.double() * 1000).float()
    tours, lens_ = [], []
    for T in (1000, 617, 2):
        t = torch.randperm(1000, generator=g)[:T].tolist()
        tours.append(t + [0] * (1000 - T))
        lens_.append(float(rl_utils.get_tour_distance(t, dm)) / 1000.0)
    out["score_x"] = xs.numpy()
    out["score_tours"] = np.array(tours, dtype=np.int64)
    out["score_tour_len"] = np.array([1000, 617, 2], dtype=np.int32)
    out["score_lengths"] = np.array(lens_, dtype=np.float32)

    path = os.path.join(OUT, "path_solver_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
