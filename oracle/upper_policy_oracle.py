"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's upper-level policy forward
(SURVEY.md section 8f rank 3): `HTSP_PPO.forward` (h_tsp.py:1272-1276) = `ActorPPO.forward`
(rl_models.py:329-336) o `IMPALAEncoder.forward` (rl_models.py:187-273) o `IMPALACNN.forward`
(rl_models.py:117-133), without torch_scatter / torch.unique: the pillar stage is written out in numpy
on the pixel index itself, the CNN and the MLP with torch.nn.functional fp32 ops.

Third-party arithmetic on this path that is absent from /root/reference: pytorch-scatter 2.0.9
(requirements.txt:159), `scatter_mean` and `scatter_max` with dim=0.  Their published semantics
(out[g] = mean / max over the rows whose index is g) are restated in `pillar_image` below and, for running
the unmodified reference, in oracle/ref_loader.py.

Pinned by tests/test_upper_policy.py against vectors recorded from the UNMODIFIED reference modules
(oracle/make_golden_policy.py -> tests/golden/upper_policy_golden.npz) and, where /root/reference is
mounted, against the live reference modules.  Only tests/, smoke() and bench.py may import this module;
the product path (h-tsp_b200/upper_policy.py) never does."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

GRID = 100                    # bev_range [0,0,1,1] / bev_pixel_size 0.01 (rl_models.py:155-158)
PIXEL = np.float32(0.01)
IN_EPS = 1e-3                 # InstanceNorm1d(eps=1e-3) (rl_models.py:178-180)


def pillar_features(state: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """state [B,N,8] f32 -> (mvf_input [B,N,12] f32, pixel id u [B*N] i64) (rl_models.py:190-222).
    u = b*GRID^2 + cx*GRID + cy is the reference's `bev_merge_coords`; rows that share u share a pillar
    (also across instances when a coordinate falls outside [0,1): the formula aliases, as it does there)."""
    B, N, C = state.shape
    pts = state.reshape(B * N, C).astype(np.float32)
    xy = pts[:, 0:2]
    idx = (xy - np.float32(0.0)) / PIXEL
    coords = np.floor(idx).astype(np.int32)
    b = np.repeat(np.arange(B, dtype=np.int64), N)
    u = b * (GRID * GRID) + coords[:, 0].astype(np.int64) * GRID + coords[:, 1].astype(np.int64)
    f_center = xy - ((coords.astype(np.float32) + np.float32(0.5)) * PIXEL + np.float32(0.0))
    # scatter_mean over equal u: fp32 running sum in row order / count
    uniq, inv, cnt = np.unique(u, return_inverse=True, return_counts=True)
    sums = np.zeros((uniq.shape[0], 2), dtype=np.float32)
    np.add.at(sums, inv, xy)
    mean = sums / np.maximum(cnt, 1).astype(np.float32)[:, None]
    f_cluster = xy - mean[inv]
    feat = np.concatenate([xy, f_center, f_cluster, pts[:, 2:]], axis=1).astype(np.float32)
    return feat.reshape(B, N, C + 4), u


def pillar_image(state: np.ndarray, w_in: np.ndarray, gamma: np.ndarray, beta: np.ndarray) -> np.ndarray:
    """-> pseudo image [B,C0,GRID,GRID] f32 indexed [b, channel, cx, cy] (rl_models.py:224-271):
    1x1 Conv1d (no bias) -> InstanceNorm1d over the N points of an instance (biased variance, eps 1e-3,
    affine) -> ReLU -> per-pillar max, zeros where no point falls.  Pillars whose id leaves
    [0, B*GRID^2) are dropped (their batch index matches no instance, rl_models.py:256)."""
    B, N, _ = state.shape
    feat, u = pillar_features(state)
    w = w_in.reshape(w_in.shape[0], -1).astype(np.float32)            # [C0,12]
    y = feat @ w.T                                                    # [B,N,C0]
    mu = y.astype(np.float64).mean(axis=1, keepdims=True)
    var = ((y.astype(np.float64) - mu) ** 2).mean(axis=1, keepdims=True)
    z = ((y - mu) / np.sqrt(var + IN_EPS)).astype(np.float32) * gamma.astype(np.float32) + beta.astype(np.float32)
    z = np.maximum(z, 0).reshape(B * N, -1)
    C0 = z.shape[1]
    img = np.zeros((B * GRID * GRID, C0), dtype=np.float32)
    ok = (u >= 0) & (u < B * GRID * GRID)
    np.maximum.at(img, u[ok], z[ok])
    return np.ascontiguousarray(img.reshape(B, GRID, GRID, C0).transpose(0, 3, 1, 2))


def impala_cnn(img: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str = "encoder.cnn.") -> torch.Tensor:
    """IMPALACNN.forward (rl_models.py:117-133): three blocks of conv3x3 -> maxpool(3,2,1) -> two residual
    blocks (ReLU, conv, ReLU, conv; + input), then ReLU(flatten) -> fc -> ReLU."""
    x = img
    i = 0
    while f"{prefix}feat_convs.{i}.0.weight" in sd:
        x = F.conv2d(x, sd[f"{prefix}feat_convs.{i}.0.weight"], sd[f"{prefix}feat_convs.{i}.0.bias"], padding=1)
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
        for blk in ("resnet1", "resnet2"):
            r = x
            t = F.conv2d(F.relu(x), sd[f"{prefix}{blk}.{i}.1.weight"], sd[f"{prefix}{blk}.{i}.1.bias"], padding=1)
            t = F.conv2d(F.relu(t), sd[f"{prefix}{blk}.{i}.3.weight"], sd[f"{prefix}{blk}.{i}.3.bias"], padding=1)
            x = t + r
        i += 1
    x = F.relu(x.flatten(1))
    return F.relu(F.linear(x, sd[f"{prefix}fc.weight"], sd[f"{prefix}fc.bias"]))


def actor(feat: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str = "actor.") -> torch.Tensor:
    """ActorPPO.forward (rl_models.py:287-301, 329-336): tanh(net(state))."""
    h = F.relu(F.linear(feat, sd[prefix + "net.0.0.weight"], sd[prefix + "net.0.0.bias"]))
    h = F.relu(F.linear(h, sd[prefix + "net.0.2.weight"], sd[prefix + "net.0.2.bias"]))
    h = F.hardswish(F.linear(h, sd[prefix + "net.1.weight"], sd[prefix + "net.1.bias"]))
    return torch.tanh(F.linear(h, sd[prefix + "net.3.weight"], sd[prefix + "net.3.bias"]))


@torch.no_grad()
def policy_forward(state: np.ndarray, sd: Dict[str, torch.Tensor]):
    """state [B,N,8] -> (image [B,C0,100,100], graph_feat [B,E], action [B,2]) as numpy f32."""
    sd = {k: torch.as_tensor(np.asarray(v), dtype=torch.float32) for k, v in sd.items()}
    img = pillar_image(np.asarray(state, dtype=np.float32), sd["encoder.input_transform.0.weight"].numpy(),
                       sd["encoder.input_transform.1.weight"].numpy(), sd["encoder.input_transform.1.bias"].numpy())
    feat = impala_cnn(torch.from_numpy(img), sd)
    act = actor(feat, sd)
    return img, feat.numpy(), act.numpy()
