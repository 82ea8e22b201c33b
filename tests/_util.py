"""Helpers shared by the parity tests."""
import math

import torch

from oracle import path_solver_oracle as po

# fp32 re-association noise of the reference's own arithmetic (measured: the same fp32 math in a
# different summation order moves the 10*tanh logits by up to 1.4e-5), so
# 1e-5 abs is not reachable by ANY independent fp32 implementation; fp32-class modes use 5e-5.
LOGIT_TOL = {"fp32": 5e-5, "x3": 1e-4, "fast": 1e-3}


def cyclic_equal(a, b):
    """True if tours a and b are the same cycle up to rotation/reflection."""
    a, b = list(a), list(b)
    if sorted(a) != sorted(b):
        return False
    n = len(a)
    i = b.index(a[0])
    fwd = [b[(i + k) % n] for k in range(n)]
    bwd = [b[(i - k) % n] for k in range(n)]
    return a == fwd or a == bwd


def tie_aware_check(sd, x, src, tgt, first, pi, tol):
    """Every decision of every row of `pi` (any implementation's greedy rollout) must be an
    argmax of the ORACLE's logits up to `tol` when the oracle is teacher-forced along that very
    tour.  Returns (max_regret, n_rows_with_nonzero_regret)."""
    x = x.cpu().float()
    pi = pi.cpu().long()
    res = po.rollout(sd, x, src.cpu().long().reshape(-1), tgt.cpu().long().reshape(-1),
                     first.cpu().long().reshape(-1), forced_pi=pi, capture_logits=True)
    assert torch.equal(res.pi, pi), "tour is not a valid rollout under the reference state machine"
    Bp, G, N = pi.shape
    srcG = src.cpu().long().reshape(Bp, 1).expand(Bp, G)
    tgtG = tgt.cpu().long().reshape(Bp, 1).expand(Bp, G)
    # replay positions: the action of step t sits at tour index `count` before the step
    count = torch.ones(Bp, G, dtype=torch.long)
    f = pi[:, :, 0]
    count = count + ((f == srcG) | (f == tgtG)).long()
    regret = torch.zeros(Bp, G)
    for t in range(N - 2):
        a = pi.gather(2, count[..., None]).squeeze(2)
        z = res.logits[t]
        chosen = z.gather(2, a[..., None]).squeeze(2)
        regret = torch.maximum(regret, z.max(dim=2).values - chosen)
        count = count + 1 + ((a == srcG) | (a == tgtG)).long()
    assert bool((count == N).all())
    worst = float(regret.max())
    assert worst <= tol, f"greedy decision off by {worst} (> {tol}) from the oracle's argmax"
    return worst, int((regret > 0).sum()), res


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(((a - b).abs() / b.abs().clamp_min(1e-12)).max())
