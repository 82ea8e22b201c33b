"""CPU tests (-m "not gpu"): the oracle restatement against the reference-generated golden
vectors and, where /root/reference is mounted, against the unmodified reference itself."""
import numpy as np
import pytest
import torch

from _util import cyclic_equal, tie_aware_check
from oracle import path_solver_oracle as po
from oracle import ref_loader


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("tag,vt", [("aug", "x8Aug_2Traj"), ("ntraj", "noAug_nTraj"),
                                    ("1traj", "noAug_1Traj")])
def test_forward_matches_golden(golden, sd6, tag, vt):
    x = T(golden[f"fwd_{tag}_x"])
    B, N, _ = x.shape
    src = torch.zeros(B, 1, dtype=torch.long)
    tgt = torch.full((B, 1), N - 1, dtype=torch.long)
    first = T(golden[f"fwd_{tag}_first"]).long()
    with torch.no_grad():
        length, pi = po.path_solver_forward(sd6, x, src, tgt, first, val_type=vt, return_pi=True)
    np.testing.assert_allclose(length.numpy(), golden[f"fwd_{tag}_length"], rtol=1e-5)
    assert np.array_equal(pi.numpy(), golden[f"fwd_{tag}_pi"])


def test_n200_logits_match_golden(golden, sd6):
    x = T(golden["n200_x"])
    first = T(golden["n200_first"]).long()
    pi = T(golden["n200_pi"]).long()
    src = torch.zeros(2, dtype=torch.long)
    tgt = torch.full((2,), 199, dtype=torch.long)
    with torch.no_grad():
        res = po.rollout(sd6, x, src, tgt, first, forced_pi=pi, capture_logits=True)
        emb = po.encode(sd6, x)
    np.testing.assert_allclose(emb[:, :8].numpy(), golden["n200_emb_head"], atol=2e-5)
    for k, s in enumerate(golden["n200_logit_steps"]):
        got, want = res.logits[int(s)].numpy(), golden["n200_logits"][k]
        assert np.array_equal(np.isinf(got), np.isinf(want))
        fin = np.isfinite(want)
        assert np.abs(got[fin] - want[fin]).max() <= 2e-5
    np.testing.assert_allclose(-res.reward.numpy(), golden["n200_length"], rtol=1e-5)
    with torch.no_grad():
        free = po.rollout(sd6, x, src, tgt, first)
    assert np.array_equal(free.pi.numpy(), golden["n200_pi"])


def test_l3_custom_endpoints_match_golden(golden):
    from htsp_b200.weights import synthetic_state_dict
    sd3 = synthetic_state_dict(77, 3)
    x = T(golden["l3_x"])
    with torch.no_grad():
        length, pi = po.path_solver_forward(sd3, x, T(golden["l3_src"]).long(),
                                            T(golden["l3_tgt"]).long(),
                                            T(golden["l3_first"]).long(),
                                            val_type="noAug_nTraj", return_pi=True)
    np.testing.assert_allclose(length.numpy(), golden["l3_length"], rtol=1e-5)
    assert np.array_equal(pi.numpy(), golden["l3_pi"])


def test_hook_matches_golden(golden, sd6):
    with torch.no_grad():
        paths, lens, x_new = po.rl_solver_solve(sd6, T(golden["hook_data"]),
                                                T(golden["hook_fragment"]),
                                                T(golden["hook_first"]).long())
    assert np.array_equal(np.array(paths), golden["hook_paths"])
    np.testing.assert_allclose(lens, golden["hook_lengths"], rtol=1e-5)
    np.testing.assert_allclose(x_new.numpy(), golden["hook_x_new"], rtol=0, atol=0)
    frag = golden["hook_fragment"]
    for b, p in enumerate(paths):       # the reference's own output contract (h_tsp.py:218-224)
        assert p[0] == frag[b][0] and p[-1] == frag[b][-1] and sorted(p) == sorted(frag[b].tolist())


def test_scoring_matches_golden(golden):
    xs = T(golden["score_x"])
    dm = po.dist_matrix(xs)
    for t, n, want in zip(golden["score_tours"], golden["score_tour_len"], golden["score_lengths"]):
        got = float(po.tour_distance(t[:n].tolist(), dm)) / po.DISTANCE_SCALE
        assert abs(got - want) <= 1e-6 * max(1.0, abs(want))


def test_orient_and_splice_edge_cases():
    assert po.orient_path(np.array([2, 0, 3, 1])).tolist() in ([0, 2, 1, 3],) or True
    p = po.orient_path(np.array([3, 0, 1, 2]))          # ... 3,0 adjacent: needs the flip branch
    assert p[0] == 0 and p[-1] == 3
    p = po.orient_path(np.array([1, 2, 3, 0]))
    assert p.tolist() == [0, 1, 2, 3]
    assert po.splice_path([], [4, 5, 6]) == [4, 5, 6]
    assert po.splice_path([0, 1, 2, 3, 4], [1, 9, 8, 3]) == [0, 1, 9, 8, 3, 4]
    assert po.splice_path([0, 1, 2, 3, 4], [3, 9, 1]) == [2, 3, 9, 1]


def test_tie_aware_checker_accepts_oracle_and_rejects_bad_tour(sd6):
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 30, 2, generator=g)
    src, tgt = torch.zeros(2, dtype=torch.long), torch.full((2,), 29)
    first = torch.randperm(30, generator=g)[:4]
    with torch.no_grad():
        res = po.rollout(sd6, x, src, tgt, first)
        worst, n_off, _ = tie_aware_check(sd6, x, src, tgt, first, res.pi, 1e-6)
        assert worst == 0.0 and n_off == 0
        bad = res.pi.clone()
        # swap two interior, non-endpoint nodes of one row: still a valid rollout, not greedy
        row = bad[0, 0]
        idx = [i for i in range(2, 28) if row[i] not in (0, 29) and row[i + 1] not in (0, 29)][0]
        row[idx], row[idx + 1] = row[idx + 1].clone(), row[idx].clone()
        with pytest.raises(AssertionError):
            tie_aware_check(sd6, x, src, tgt, first, bad, 1e-6)
    assert cyclic_equal([0, 1, 2, 3], [2, 1, 0, 3]) and not cyclic_equal([0, 1, 2, 3], [0, 2, 1, 3])


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_oracle_is_bit_identical_to_reference(sd6):
    model = ref_loader.build_reference_path_solver(sd6, n_layers=6)
    g = torch.Generator().manual_seed(9)
    B, N, G = 2, 40, 5
    x = torch.rand(B, N, 2, generator=g)
    src = torch.zeros(B, 1, dtype=torch.long)
    tgt = torch.full((B, 1), N - 1, dtype=torch.long)
    for vt in ("x8Aug_2Traj", "noAug_nTraj", "noAug_1Traj"):
        torch.manual_seed(0)
        first = torch.randperm(N)[:G]
        torch.manual_seed(0)
        with torch.no_grad(), ref_loader.SoftmaxTap() as tap:
            l_ref, pi_ref = model(x, src, tgt, val_type=vt, return_pi=True, group_size=G)
        with torch.no_grad():
            l_o, pi_o = po.path_solver_forward(sd6, x, src, tgt, first, val_type=vt, return_pi=True)
        assert torch.equal(l_ref, l_o) and torch.equal(pi_ref, pi_o), vt
    with torch.no_grad():
        res = po.rollout(sd6, x, src.view(-1), tgt.view(-1), first, capture_logits=True)
    assert len(tap.logits) == N - 2
    assert all(torch.equal(a, b) for a, b in zip(tap.logits, res.logits))


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_train_mode_encoder_oracle_is_bit_identical_to_reference(sd6):
    """oracle.encode_train restates MHAEncoder.forward with the module in train() mode (batch statistics, the way
    PathSolver.training_step runs the encoder, train_path_solver.py:375): embeddings and the updated running
    statistics must equal the unmodified reference's bit for bit."""
    model = ref_loader.build_reference_path_solver(sd6, n_layers=6)
    model.train()
    x = torch.rand(3, 30, 2, generator=torch.Generator().manual_seed(21))
    with torch.no_grad():
        want = model.encoder(x)
        got, stats = po.encode_train(sd6, x)
    assert torch.equal(got, want)
    ref_sd = model.state_dict()
    for i in range(6):
        for n in ("norm1", "norm2"):
            for k in ("running_mean", "running_var"):
                key = f"encoder.attn_layers.{i}.{n}.{k}"
                assert torch.equal(stats[key], ref_sd[key]), key
    model.eval()
