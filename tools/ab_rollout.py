"""A/B timing of the rollout kernel: decode-only rate (CUDA events around the rollout kernel, through the library's
kernel-timer hook) and a checksum of the tours, for the library named by HTSP_LIB_PATH (default: the in-tree
build).  Run it once per build and compare the lines: the same checksum means bit-identical tours."""
import argparse, hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import htsp_b200  # noqa
from htsp_b200 import B200PathSolver, _abi
from htsp_b200.weights import synthetic_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=200)
ap.add_argument("--g", type=int, default=200)
ap.add_argument("--instances", type=int, default=592)
ap.add_argument("--mode", default="x3")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--layers", type=int, default=6)
ap.add_argument("--tag", default="")
a = ap.parse_args()
s = B200PathSolver(state_dict=synthetic_state_dict(1234, a.layers), mode=a.mode, device="cuda")
g = torch.Generator().manual_seed(0)
x = torch.rand(a.instances, a.n, 2, generator=g).cuda()
src = torch.zeros(a.instances, dtype=torch.long); tgt = torch.full((a.instances,), a.n - 1)
first = torch.randperm(a.n, generator=g)[:a.g]
emb = s.encode(x)
lib = _abi.lib()
ms = []
for r in range(a.reps + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); e1.record()        # materialise the cudaEvent_t handles; the library re-records them
    lib.htsp_set_kernel_timer(e0.cuda_event, e1.cuda_event)
    pi, reward, _ = s.rollout(x, emb, src, tgt, first)
    torch.cuda.synchronize()
    if r > 0:
        ms.append(e0.elapsed_time(e1))
h = hashlib.sha1(pi.cpu().numpy().tobytes()).hexdigest()[:16]
best = min(ms)
print(json.dumps({"tag": a.tag, "lib": os.path.basename(_abi.LIB_PATH), "mode": a.mode, "N": a.n, "G": a.g,
                  "instances": a.instances, "rollout_ms": [round(m, 3) for m in ms],
                  "instances_per_s": round(a.instances / best * 1e3, 1), "tours_sha1": h,
                  "mean_reward": round(float(reward.mean()), 6)}))
