"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain torch ops, fp32 by default) of the
H-TSP lower-level path-solver hot path.  Nothing in the product package may import this
file; it is the checker used by `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py`.

Parity pin: this restatement is checked against the UNMODIFIED reference imported from
/root/reference (oracle/ref_loader.py) by tests/test_oracle_vs_reference.py (runs only where
the reference tree is mounted) and against the committed golden vectors in tests/golden/
that were produced by the reference itself (oracle/make_golden.py).  The reference ships no
tests or golden vectors of its own for this path.

Every function cites the reference lines (relative to /root/reference) it restates.  The op
order inside each function follows the reference so that on CPU/fp32 the results are
bit-identical to it; the *structure* differs (pure functions over a state_dict, a per-row
tour buffer instead of cat + unique_consecutive).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
N_HEADS = 8
BN_EPS = 1e-5
TANH_CLIP = 10.0


# --------------------------------------------------------------------------------------
# a5: 8-fold augmentation                         rl4cop/utils/utils.py:37-56
# --------------------------------------------------------------------------------------
def augment_8fold(xy: Tensor) -> Tensor:
    """[B,N,2] -> [8B,N,2], block-major: block k holds isometry k of every instance."""
    x, y = xy[:, :, [0]], xy[:, :, [1]]
    blocks = [(x, y), (1 - x, y), (x, 1 - y), (1 - x, 1 - y),
              (y, x), (1 - y, x), (y, 1 - x), (1 - y, 1 - x)]
    return torch.cat([torch.cat(b, dim=2) for b in blocks], dim=0)


# --------------------------------------------------------------------------------------
# a10: attention helpers                          rl4cop/utils/utils.py:12-34
# --------------------------------------------------------------------------------------
def split_heads(t: Tensor, n_heads: int = N_HEADS) -> Tensor:
    """[B,n,H] -> [B,heads,n,dk]                  (utils.py:32-34)"""
    return t.reshape(t.size(0), t.size(1), n_heads, -1).transpose(1, 2)


def attention(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """softmax(q k^T / sqrt(dk) + mask) v, softmax in fp32   (utils.py:12-29)"""
    score = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(q.size(-1))
    if mask is not None:
        score = score + mask[:, None, :, :].expand_as(score)
    p = F.softmax(score.float(), dim=3, dtype=torch.float32).type_as(score)
    out = torch.matmul(p, v).transpose(1, 2)
    return out.reshape(q.size(0), q.size(2), q.size(1) * q.size(3))


# --------------------------------------------------------------------------------------
# a9: encoder                                     rl4cop/models/models.py:12-60
# --------------------------------------------------------------------------------------
def n_layers_of(sd: Dict[str, Tensor]) -> int:
    n = 0
    while f"encoder.attn_layers.{n}.Wq.weight" in sd:
        n += 1
    return n


def _bn_eval(x: Tensor, sd, prefix: str) -> Tensor:
    """BatchNorm1d in eval mode over the flattened [B*N, H] view (models.py:35,37)."""
    shp = x.shape
    y = F.batch_norm(x.reshape(-1, shp[-1]), sd[prefix + ".running_mean"],
                     sd[prefix + ".running_var"], sd[prefix + ".weight"],
                     sd[prefix + ".bias"], False, 0.0, BN_EPS)
    return y.view(*shp)


def encoder_layer(sd, i: int, x: Tensor) -> Tensor:
    """One MHAEncoderLayer.forward (models.py:30-38)."""
    p = f"encoder.attn_layers.{i}."
    q = split_heads(F.linear(x, sd[p + "Wq.weight"]))
    k = split_heads(F.linear(x, sd[p + "Wk.weight"]))
    v = split_heads(F.linear(x, sd[p + "Wv.weight"]))
    x = x + F.linear(attention(q, k, v), sd[p + "multi_head_combine.weight"],
                     sd[p + "multi_head_combine.bias"])
    x = _bn_eval(x, sd, p + "norm1")
    h = F.linear(x, sd[p + "feed_forward.0.weight"], sd[p + "feed_forward.0.bias"])
    h = F.linear(F.relu(h), sd[p + "feed_forward.2.weight"], sd[p + "feed_forward.2.bias"])
    x = x + h
    x = _bn_eval(x, sd, p + "norm2")
    return x


def _bn_train(x: Tensor, sd, prefix: str, stats: dict) -> Tensor:
    """torch.nn.BatchNorm1d in train() mode on x.view(-1, C) (models.py:35,37): batch statistics; `stats` receives
    the updated running_mean / running_var (momentum 0.1, unbiased variance) under the state_dict keys."""
    flat = x.reshape(-1, x.size(-1))
    rm, rv = sd[prefix + ".running_mean"].clone(), sd[prefix + ".running_var"].clone()
    y = F.batch_norm(flat, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], True, 0.1, 1e-5)
    stats[prefix + ".running_mean"], stats[prefix + ".running_var"] = rm, rv
    stats[prefix + ".batch_mean"] = flat.mean(dim=0)
    stats[prefix + ".batch_var"] = flat.var(dim=0, unbiased=False)
    return y.view(*x.size())


def encode_train(sd, x: Tensor):
    """MHAEncoder.forward with the module in train() mode (as PathSolver.training_step runs it,
    train_path_solver.py:375): returns (embeddings, dict of updated running statistics + batch statistics)."""
    stats = {}
    h = F.linear(x, sd["encoder.init_projection_layer.weight"], sd["encoder.init_projection_layer.bias"])
    for i in range(n_layers_of(sd)):
        p = f"encoder.attn_layers.{i}."
        q = split_heads(F.linear(h, sd[p + "Wq.weight"]))
        k = split_heads(F.linear(h, sd[p + "Wk.weight"]))
        v = split_heads(F.linear(h, sd[p + "Wv.weight"]))
        h = h + F.linear(attention(q, k, v), sd[p + "multi_head_combine.weight"], sd[p + "multi_head_combine.bias"])
        h = _bn_train(h, sd, p + "norm1", stats)
        f = F.linear(h, sd[p + "feed_forward.0.weight"], sd[p + "feed_forward.0.bias"])
        f = F.linear(F.relu(f), sd[p + "feed_forward.2.weight"], sd[p + "feed_forward.2.bias"])
        h = _bn_train(h + f, sd, p + "norm2", stats)
    return h, stats


def encode(sd, x: Tensor) -> Tensor:
    """MHAEncoder.forward (models.py:55-60): coords [B',N,2] -> embeddings [B',N,H]."""
    h = F.linear(x, sd["encoder.init_projection_layer.weight"],
                 sd["encoder.init_projection_layer.bias"])
    for i in range(n_layers_of(sd)):
        h = encoder_layer(sd, i, h)
    return h


# --------------------------------------------------------------------------------------
# a11/a12: decoder                                rl4cop/models/models.py:360-446
# --------------------------------------------------------------------------------------
class DecoderCache:
    """What PathDecoder.reset caches (models.py:375-402)."""

    def __init__(self, sd, emb: Tensor, src: Tensor, tgt: Tensor, first: Tensor):
        B, N, H = emb.shape
        G = first.size(1)
        self.sd, self.emb = sd, emb

        def node_q(w: str, idx: Tensor) -> Tensor:
            e = emb.gather(1, idx.view(B, G, 1).expand(B, G, H))
            return split_heads(F.linear(e, sd[w]))

        self.q_graph = split_heads(F.linear(emb.mean(dim=1, keepdim=True),
                                            sd["decoder.Wq_graph.weight"]))
        self.q_source = node_q("decoder.Wq_source.weight", src)
        self.q_target = node_q("decoder.Wq_target.weight", tgt)
        self.q_first = node_q("decoder.Wq_first.weight", first)
        self.k = split_heads(F.linear(emb, sd["decoder.Wk.weight"]))
        self.v = split_heads(F.linear(emb, sd["decoder.Wv.weight"]))
        self.logit_k = emb.transpose(1, 2)


def decoder_logits(c: DecoderCache, last: Tensor, ninf_mask: Tensor) -> Tensor:
    """PathDecoder.forward up to (not including) the final softmax (models.py:404-440):
    returns the masked, 10*tanh-clipped pointer logits [B',G,N]."""
    B, N, H = c.emb.shape
    G = last.size(1)
    e_last = c.emb.gather(1, last.view(B, G, 1).expand(-1, -1, H))
    q_last = split_heads(F.linear(e_last, c.sd["decoder.Wq_last.weight"]))
    q = c.q_graph + c.q_source + c.q_target + c.q_first + q_last
    glimpse = attention(q, c.k, c.v, ninf_mask)
    final_q = F.linear(glimpse, c.sd["decoder.multi_head_combine.weight"],
                       c.sd["decoder.multi_head_combine.bias"])
    score = torch.matmul(final_q, c.logit_k) / math.sqrt(H)
    return TANH_CLIP * torch.tanh(score) + ninf_mask


# --------------------------------------------------------------------------------------
# a6/a7/a8/a13/a14: rollout                        rl4cop/train_path_solver.py:23-149,240-288
# --------------------------------------------------------------------------------------
class RolloutResult:
    def __init__(self, pi, reward, logits):
        self.pi = pi            # [B',G,N] int64, cycle order starting at first_action[g]
        self.reward = reward    # [B',G] f32 = -(closed cycle length - |x[src]-x[tgt]|)
        self.logits = logits    # list over the N-2 decode steps of [B',G,N] (or None)


def _path_reward(x: Tensor, pi: Tensor, src: Tensor, tgt: Tensor) -> Tensor:
    """Env._get_path_distance / _get_edge_length (train_path_solver.py:109-149)."""
    B, G, N = pi.shape
    C = x.size(2)
    seq = x[:, None, :, :].expand(B, G, N, C)
    ordered = seq.gather(2, pi.unsqueeze(3).expand(B, G, N, C))
    delta = ordered - ordered.roll(dims=2, shifts=-1)
    tour_len = (delta ** 2).sum(3).sqrt().sum(2)
    ends = torch.cat([src[..., None, None].expand(B, G, 1, C),
                      tgt[..., None, None].expand(B, G, 1, C)], dim=2)
    e = seq.gather(2, ends)
    fixed = ((e - e.roll(dims=2, shifts=-1))[:, :, :-1, :] ** 2).sum(3).sqrt().sum(2)
    return -(tour_len - fixed)


def rollout(sd, x: Tensor, src: Tensor, tgt: Tensor, first_action: Tensor,
            forced_pi: Optional[Tensor] = None, capture_logits: bool = False,
            emb: Optional[Tensor] = None) -> RolloutResult:
    """Greedy POMO rollout of every (instance b', start g) row.

    x [B',N,2]; src,tgt [B'] int64; first_action [G] int64 (the reference draws it with
    torch.randperm(N)[:G] at train_path_solver.py:256 -- here it is an argument so that both
    sides of a parity test see the same starts).  State machine = SURVEY.md section 3.5
    (GroupState.move_to, train_path_solver.py:45-72): visiting src immediately visits tgt
    (and vice versa) and the partner becomes the current node.

    forced_pi [B',G,N]: teacher forcing -- the action of each step is read from this tour
    instead of argmax (used to compare per-step logits without divergence).
    """
    Bp, N, _ = x.shape
    G = first_action.numel()
    dev = x.device
    srcG = src.view(Bp, 1).expand(Bp, G)
    tgtG = tgt.view(Bp, 1).expand(Bp, G)
    first = first_action.view(1, G).expand(Bp, G)
    if emb is None:
        emb = encode(sd, x)

    tour = torch.zeros(Bp, G, N, dtype=torch.long, device=dev)
    count = torch.zeros(Bp, G, dtype=torch.long, device=dev)
    mask = torch.zeros(Bp, G, N, dtype=x.dtype, device=dev)

    def visit(a: Tensor) -> Tensor:
        nonlocal count
        tour.scatter_(2, count[..., None], a[..., None])
        mask.scatter_(2, a[..., None], -math.inf)
        count = count + 1
        partner = torch.where(a == srcG, tgtG, torch.where(a == tgtG, srcG, a))
        dbl = partner != a
        # rows with a coupled visit append the partner as well (train_path_solver.py:49-51,60-66)
        pos = torch.where(dbl, count, torch.zeros_like(count))
        old = tour.gather(2, pos[..., None]).squeeze(2)
        tour.scatter_(2, pos[..., None], torch.where(dbl, partner, old)[..., None])
        mask.scatter_(2, partner[..., None], -math.inf)
        count = count + dbl.long()
        return partner

    cur = visit(first)                                        # step 0: no decoder call (:256-257)
    cache = DecoderCache(sd, emb, srcG, tgtG, first)          # reset AFTER step 0 (:258-260)
    logits_out: List[Tensor] = []
    for _ in range(N - 2):                                    # (:261-278)
        z = decoder_logits(cache, cur, mask)
        if capture_logits:
            logits_out.append(z.clone())
        if forced_pi is not None:
            a = forced_pi.gather(2, count[..., None]).squeeze(2)
        else:
            probs = F.softmax(z.float(), dim=2, dtype=torch.float32).type_as(z)
            a = probs.argmax(dim=2)                           # greedy (:263-264)
        cur = visit(a)
    assert bool((count == N).all()), "every row must end with all N nodes (:285-287)"
    reward = _path_reward(x, tour, srcG, tgtG)
    return RolloutResult(tour, reward, logits_out if capture_logits else None)


# --------------------------------------------------------------------------------------
# a15: best-of selection + full forward           rl4cop/train_path_solver.py:224-311
# --------------------------------------------------------------------------------------
def select_best(reward: Tensor, pi: Tensor, val_type: str) -> Tuple[Tensor, Tensor]:
    """(:290-306) -> (max_reward, best_pi) before the final squeeze."""
    Bp, G, N = pi.shape
    if val_type == "noAug_1Traj":
        return reward, pi
    if val_type == "noAug_nTraj":
        best, i1 = reward.max(dim=1)
        return best, pi.gather(1, i1.reshape(Bp, 1, 1).repeat(1, 1, N))
    B = round(Bp / 8)
    r = reward.reshape(8, B, G)
    best, i2 = r.max(dim=2)
    best, i0 = best.max(dim=0)
    p = pi.reshape(8, B, G, N)
    i0 = i0.reshape(1, B, 1, 1)
    i2 = i2.reshape(8, B, 1, 1).gather(0, i0)
    bp = p.gather(0, i0.repeat(1, 1, G, N)).gather(2, i2.repeat(1, 1, 1, N))
    return best, bp


def path_solver_forward(sd, batch: Tensor, source_nodes: Tensor, target_nodes: Tensor,
                        first_action: Tensor, val_type: str = "x8Aug_2Traj",
                        return_pi: bool = False):
    """PathSolver.forward (train_path_solver.py:224-311) with greedy=True and the POMO
    starts passed in (group_size = first_action.numel())."""
    if "x8Aug" in val_type:
        batch = augment_8fold(batch)
        source_nodes = torch.repeat_interleave(source_nodes, 8, dim=0)
        target_nodes = torch.repeat_interleave(target_nodes, 8, dim=0)
    res = rollout(sd, batch, source_nodes.reshape(-1), target_nodes.reshape(-1), first_action)
    best, best_pi = select_best(res.reward, res.pi, val_type)
    if return_pi:
        return -best, best_pi.squeeze()
    return -best


# --------------------------------------------------------------------------------------
# a2/a3/a16: the RLSolver hook                    h_tsp.py:165-228
# --------------------------------------------------------------------------------------
def extract_normalize(data: Tensor, fragment: Tensor) -> Tuple[Tensor, Tensor]:
    """gather + min/max normalisation into [0.05,0.95]^2 (h_tsp.py:176-191) -> (x_new, s)."""
    x = torch.gather(data, 1, fragment[..., None].expand(-1, -1, data.shape[-1]))
    x1_min = x[..., 0].min(dim=-1, keepdim=True)[0]
    x2_min = x[..., 1].min(dim=-1, keepdim=True)[0]
    x1_max = x[..., 0].max(dim=-1, keepdim=True)[0]
    x2_max = x[..., 1].max(dim=-1, keepdim=True)[0]
    s = 0.9 / torch.maximum(x1_max - x1_min, x2_max - x2_min)
    x_new = torch.empty_like(x)
    x_new[..., 0] = s * (x[..., 0] - x1_min) + 0.05
    x_new[..., 1] = s * (x[..., 1] - x2_min) + 0.05
    return x_new, s


def orient_path(path: np.ndarray) -> np.ndarray:
    """Rotate/flip a cyclic tour so it starts at local node 0 and ends at N-1 (h_tsp.py:214-224)."""
    n = path.shape[0]
    zero_idx = np.nonzero(path == 0)[0][0]
    path = np.roll(path, -zero_idx)
    if path[-1] != n - 1:
        path = np.flip(np.roll(path, -1))
    assert path[0] == 0 and path[-1] == n - 1 and np.unique(path).shape[0] == n
    return path


def rl_solver_solve(sd, data: Tensor, fragment: Tensor, first_action: Tensor,
                    val_type: str = "x8Aug_2Traj"):
    """RLSolver.solve (h_tsp.py:170-228) -> (paths as global ids, lengths, x_new)."""
    x_new, s = extract_normalize(data, fragment)
    B, N = fragment.shape
    src = torch.zeros(B, 1, dtype=torch.long, device=data.device)
    tgt = torch.full((B, 1), N - 1, dtype=torch.long, device=data.device)
    lengths, paths = path_solver_forward(sd, x_new.float(), src, tgt, first_action,
                                         val_type=val_type, return_pi=True)
    lengths = (lengths / s.squeeze()).cpu().numpy()
    paths = paths.cpu().numpy()
    frag = fragment.cpu().numpy()
    if B == 1:
        paths = [paths]
    out = [[int(frag[i][j]) for j in orient_path(p)] for i, p in enumerate(paths)]
    return out, lengths, x_new


# --------------------------------------------------------------------------------------
# a17: scoring                                    h_tsp.py:449-453,586-589; rl_utils.py:87-102
# --------------------------------------------------------------------------------------
DISTANCE_SCALE = 1000


def dist_matrix(x: Tensor) -> Tensor:
    """fp64 cdist of 1000*x rounded to fp32 (h_tsp.py:449-453)."""
    xs = x.type(torch.float64) * DISTANCE_SCALE
    return torch.cdist(xs, xs).type(torch.float32)


def tour_distance(tour: List[int], dm: Tensor) -> Tensor:
    """get_tour_distance (rl_utils.py:87-102): fp32 sum of dm[t_i, t_{i+1}] over the closed tour."""
    t = torch.tensor(tour, dtype=torch.int64, device=dm.device)
    nxt = torch.tensor(tour[1:] + tour[0:1], dtype=torch.int64, device=dm.device)
    return dm[t, nxt].sum()


def splice_path(current_tour: List[int], new_path: List[int]) -> List[int]:
    """LargeState.move_to list splice (h_tsp.py:288-307)."""
    if len(current_tour) == 0:
        return list(new_path)
    s, e = current_tour.index(new_path[0]), current_tour.index(new_path[-1])
    assert s != e
    if e > s:
        return current_tour[:s] + list(new_path) + current_tour[e + 1:]
    return current_tour[e + 1: s] + list(new_path)
