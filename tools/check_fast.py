"""Quick parity probe of a rollout mode against the CPU oracle (teacher-forced logits of every step, then a free-running
tie-aware check) on a few shapes; used while developing kernels (the judged tests are tests/test_parity_gpu.py)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import htsp_b200  # noqa
from htsp_b200 import B200PathSolver
from htsp_b200.weights import synthetic_state_dict
from oracle import path_solver_oracle as po
from _util import tie_aware_check

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="fast")
ap.add_argument("--shapes", default="2x50x7,1x200x33,2x100x100,1x208x208,2x130x130,1x150x129,1x64x1")
ap.add_argument("--seed", type=int, default=1234)
ap.add_argument("--layers", type=int, default=6)
a = ap.parse_args()
sd = synthetic_state_dict(a.seed, a.layers)
s = B200PathSolver(state_dict=sd, mode=a.mode, device="cuda")
for shp in a.shapes.split(","):
    Bp, N, G = (int(v) for v in shp.split("x"))
    g = torch.Generator().manual_seed(N + G)
    x = torch.rand(Bp, N, 2, generator=g)
    src = torch.randint(0, N, (Bp,), generator=g)
    tgt = (src + 1 + torch.randint(0, N - 1, (Bp,), generator=g)) % N
    first = torch.randperm(N, generator=g)[:G]
    with torch.no_grad():
        ref = po.rollout(sd, x, src, tgt, first, capture_logits=True)
    xc = x.cuda()
    emb = s.encode(xc)
    pi, reward, logits = s.rollout(xc, emb, src, tgt, first, forced_pi=ref.pi, want_logits=True)
    torch.cuda.synchronize()
    want = torch.stack(ref.logits, dim=2)
    got = logits.cpu()
    fin = torch.isfinite(want)
    out = {"mode": a.mode, "shape": shp, "pi_equal": bool(torch.equal(pi.cpu().long(), ref.pi)),
           "inf_pattern_equal": bool(torch.equal(torch.isinf(got), torch.isinf(want))),
           "max_logit_err": float((got[fin] - want[fin]).abs().max()),
           "reward_rel_err": float(((reward.cpu() - ref.reward).abs() / ref.reward.abs()).max())}
    pi2, reward2, _ = s.rollout(xc, emb, src, tgt, first)
    try:
        with torch.no_grad():
            worst, n_off, forced = tie_aware_check(sd, x, src, tgt, first, pi2.cpu().long(), 2e-3)
        out["free_run_worst_regret"] = worst
        out["free_run_reward_rel_err"] = float(((reward2.cpu() - forced.reward).abs() / forced.reward.abs()).max())
    except AssertionError as e:
        out["free_run_error"] = str(e)[:200]
    print(json.dumps(out), flush=True)
