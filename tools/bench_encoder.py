"""Encoder-only timing (K2: init projection + L layers + nothing else) with CUDA events, fused feed-forward kernel
against the two-GEMM path (HTSP_FF_FUSED=0), and a bit-identity check of the embeddings.
    python tools/bench_encoder.py [--instances 4096] [--n 200] [--reps 5]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import htsp_b200  # noqa
from htsp_b200 import B200PathSolver
from htsp_b200.weights import synthetic_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--instances", type=int, default=4096)
ap.add_argument("--n", type=int, default=200)
ap.add_argument("--layers", type=int, default=6)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--mode", default="fast")
a = ap.parse_args()
s = B200PathSolver(state_dict=synthetic_state_dict(1234, a.layers), mode=a.mode, device="cuda")
x = torch.rand(a.instances, a.n, 2, generator=torch.Generator().manual_seed(0)).cuda()
out, embs = {}, {}
for fused in ("0", "1"):
    os.environ["HTSP_FF_FUSED"] = fused
    embs[fused] = s.encode(x).clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        s.encode(x)
    e1.record()
    torch.cuda.synchronize()
    out["fused_ms" if fused == "1" else "two_gemm_ms"] = e0.elapsed_time(e1) / a.reps
print(json.dumps({"instances": a.instances, "N": a.n, "layers": a.layers, **out,
                  "bit_identical": bool(torch.equal(embs["0"], embs["1"]))}))
