"""Per-phase cycle counts of the tcgen05 rollout kernel (DEBUG instantiation, clock64 timers of instance 0 while
`--instances` CTAs keep the other SMs busy).  Rows: softmax warps sw = 0..15 (tile, key half, lane quarter) with
phases [wait S, softmax, wait O + drain, wait U, arg-max, state + gather]; issuers 16, 17 with phases
[ring wait, idle poll for P, P V / S turns, wait Q, wait O staged, logits issue].  Cycles per decode step."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import htsp_b200  # noqa
from htsp_b200 import B200PathSolver
from htsp_b200.weights import synthetic_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200)
ap.add_argument("--g", type=int, default=200)
ap.add_argument("--instances", type=int, default=148)
ap.add_argument("--mode", default="x3")
a = ap.parse_args()
s = B200PathSolver(state_dict=synthetic_state_dict(1234, 6), mode=a.mode, device="cuda")
g = torch.Generator().manual_seed(0)
x = torch.rand(a.instances, a.n, 2, generator=g).cuda()
src = torch.zeros(a.instances, dtype=torch.long); tgt = torch.full((a.instances,), a.n - 1)
first = torch.randperm(a.n, generator=g)[:a.g]
emb = s.encode(x)
s.rollout_debug(x, emb, src, tgt, first, None, 3)
cyc = (s.last_debug_cycles.cpu() / (a.n - 2)).round().long().tolist()
if a.mode == "fast":     # decode_tc2.cu: 24 softmax warps (tile, key part, lane quarter), issuers in rows 24, 25
    out = {"mode": a.mode, "N": a.n, "G": a.g, "instances": a.instances,
           "softmax_warps": {f"t{w // 12}.part{(w % 12) >> 2}.q{w & 3}": cyc[w] for w in range(24)},
           "issuers": {f"tile{t}": cyc[24 + t] for t in range(2)}}
else:
    out = {"mode": a.mode, "N": a.n, "G": a.g, "instances": a.instances,
           "softmax_warps": {f"t{w >> 3}.half{(w >> 2) & 1}.q{w & 3}": cyc[w] for w in range(16)},
           "issuers": {f"tile{t}": cyc[16 + t] for t in range(2)}}
print(json.dumps(out))
