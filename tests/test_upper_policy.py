"""Upper-level policy forward (h_tsp.py:1272-1276, rl_models.py:117-336):
oracle vs vectors recorded from the unmodified reference (CPU), kernels vs oracle / golden through the
C ABI (GPU).  Tolerances: everything is fp32 on both sides with different summation orders (direct
convolution vs the reference's GEMM-based one): features within 2e-5 of the largest feature, actions
(tanh outputs) within 2e-5 abs, pseudo-image within 1e-5 abs; the measured values are ~10x smaller."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import upper_policy_oracle as uo

from htsp_b200 import upper_policy as pol

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "upper_policy_golden.npz")
CASES = ("n300", "n1000", "cluster")
FEAT_TOL, ACT_TOL, IMG_TOL = 2e-5, 2e-5, 1e-5


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def sd(golden):
    return pol.synthetic_policy_state_dict(int(golden["seed"]))


def _check_against_golden(g, tag, img, feat, act):
    assert int((img != 0).sum()) == int(g[f"{tag}_img_nonzero"])
    ref_sum = g[f"{tag}_img_chan_sum"]
    assert np.abs(img.astype(np.float64).sum(axis=(2, 3)) - ref_sum).max() <= 1e-6 * np.abs(ref_sum).max() + 1e-3
    assert np.abs(img[:, :, ::7, :].max(axis=3) - g[f"{tag}_img_rows"]).max() <= IMG_TOL
    gf = g[f"{tag}_feat"]
    assert np.abs(feat - gf).max() <= FEAT_TOL * np.abs(gf).max()
    assert np.abs(act - g[f"{tag}_action"]).max() <= ACT_TOL


# ------------------------------------------------------------------------------------------ CPU
def test_synthetic_weights_are_portable(golden, sd):
    cs = sum(float(v.double().abs().sum()) for v in sd.values())
    assert abs(cs - float(golden["weights_checksum"])) < 1e-6 * cs
    assert [k for k, _ in pol.expected_policy_shapes()] == list(sd.keys())
    assert pol.FLAT_DIM == 5408


@pytest.mark.parametrize("tag", CASES)
def test_oracle_matches_reference_golden(golden, sd, tag):
    img, feat, act = uo.policy_forward(golden[f"{tag}_state"], sd)
    _check_against_golden(golden, tag, img, feat, act)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_oracle_matches_live_reference(sd):
    M = ref_loader.load_rl_models_module()
    enc = M.IMPALAEncoder(input_dim=8, embedding_dim=128)
    act = M.ActorPPO(state_dim=128, mid_dim=128, action_dim=2, init_a_std_log=-1)
    # the key list the C ABI takes is exactly the reference's state_dict (minus a_std_log)
    ref_keys = ["encoder." + k for k in enc.state_dict()] + ["actor." + k for k in act.state_dict() if k != "a_std_log"]
    assert sorted(ref_keys) == sorted(k for k, _ in pol.expected_policy_shapes())
    for k, shape in pol.expected_policy_shapes():
        mod, name = (enc, k[len("encoder."):]) if k.startswith("encoder.") else (act, k[len("actor."):])
        assert tuple(mod.state_dict()[name].shape) == shape
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
    act.load_state_dict({k[len("actor."):]: v for k, v in sd.items() if k.startswith("actor.")}, strict=False)
    enc.eval(), act.eval()
    rs = np.random.RandomState(5)
    s = rs.rand(4, 700, 8).astype(np.float32)
    s[1, :40, 0:2] = s[1, 0, 0:2]                       # 40 cities in one pillar
    with torch.no_grad():
        f_ref = enc(torch.from_numpy(s))
        a_ref = act(f_ref)
    _, feat, a = uo.policy_forward(s, sd)
    assert np.abs(feat - f_ref.numpy()).max() <= FEAT_TOL * float(f_ref.abs().max())
    assert np.abs(a - a_ref.numpy()).max() <= ACT_TOL


# ------------------------------------------------------------------------------------------ GPU
def _gpu_policy(sd):
    return pol.B200UpperPolicy(sd, "cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
def test_kernels_match_reference_golden(built_lib, golden, sd, tag):
    p = _gpu_policy(sd)
    act, feat, img = p(torch.from_numpy(golden[f"{tag}_state"]).cuda(), return_features=True, return_image=True)
    _check_against_golden(golden, tag, img.cpu().numpy(), feat.cpu().numpy(), act.cpu().numpy())


def _random_states(rs, E, N, spread=1.0):
    s = rs.rand(E, N, 8).astype(np.float32)
    s[:, :, 0:2] = (0.5 + (s[:, :, 0:2] - 0.5) * spread).astype(np.float32)
    s[:, :, 2] = s[:, :, 2] < 0.6
    s[:, :, 3] = (s[:, :, 3] < 0.4) * s[:, :, 2]
    return s


@pytest.mark.gpu
@pytest.mark.parametrize("E,N,spread", [(16, 1000, 1.0), (3, 10000, 1.0), (2, 1, 1.0), (5, 333, 0.05), (1, 4097, 1.0)])
def test_kernels_match_oracle(built_lib, sd, E, N, spread):
    rs = np.random.RandomState(1000 + E + N)
    s = _random_states(rs, E, N, spread)
    p = _gpu_policy(sd)
    act, feat, img = p(torch.from_numpy(s).cuda(), return_features=True, return_image=True)
    o_img, o_feat, o_act = uo.policy_forward(s, sd)
    img = img.cpu().numpy()
    assert np.array_equal(img != 0, o_img != 0)
    assert np.abs(img - o_img).max() <= IMG_TOL
    assert np.abs(feat.cpu().numpy() - o_feat).max() <= FEAT_TOL * np.abs(o_feat).max()
    assert np.abs(act.cpu().numpy() - o_act).max() <= ACT_TOL
    # bit-reproducible: fixed-point pillar sums, integer max, fixed reduction orders
    act2, feat2 = p(torch.from_numpy(s).cuda(), return_features=True)
    assert torch.equal(act, act2) and torch.equal(feat, feat2)


@pytest.mark.gpu
def test_out_of_range_coordinates_alias_like_the_reference(built_lib, sd):
    """x = 1.0 gives pixel row 100: the reference's pillar id then points into the NEXT instance's image
    (rl_models.py:207-211, 256) and falls off the end for the last instance."""
    rs = np.random.RandomState(3)
    s = _random_states(rs, 3, 400)
    s[0, 7, 0] = 1.0
    s[2, 9, 0] = 1.0
    s[1, 11, 1] = 1.0
    p = _gpu_policy(sd)
    act, feat, img = p(torch.from_numpy(s).cuda(), return_features=True, return_image=True)
    o_img, o_feat, o_act = uo.policy_forward(s, sd)
    assert np.array_equal(img.cpu().numpy() != 0, o_img != 0)
    assert np.abs(img.cpu().numpy() - o_img).max() <= IMG_TOL
    assert np.abs(act.cpu().numpy() - o_act).max() <= ACT_TOL


@pytest.mark.gpu
def test_weight_validation(built_lib, sd):
    bad = dict(sd)
    bad.pop("actor.net.1.bias")
    with pytest.raises(KeyError):
        pol.B200UpperPolicy(bad, "cuda:0")
    bad = dict(sd)
    bad["encoder.cnn.fc.weight"] = bad["encoder.cnn.fc.weight"][:, :100]
    with pytest.raises(ValueError):
        pol.B200UpperPolicy(bad, "cuda:0")
    p = _gpu_policy(sd)
    with pytest.raises(AssertionError):
        p(torch.zeros(2, 10, 7, device="cuda"))


@pytest.mark.gpu
def test_policy_drives_the_device_environment(built_lib, sd):
    """evaluate.py:82-89 on the device: policy -> fragment -> path solver -> splice, until every
    environment is done; tours are permutations and the final length matches a recomputation."""
    from htsp_b200 import B200PathSolver, B200RLSolver
    from htsp_b200.vec_env import B200VecEnv
    from htsp_b200.weights import synthetic_state_dict
    torch.manual_seed(11)
    E, N, L = 4, 400, 64
    x = torch.rand(E, N, 2, device="cuda")
    solver = B200PathSolver(state_dict=synthetic_state_dict(1234, 3), mode="x3", device="cuda", graph_size=L)
    hook = B200RLSolver(solver, sample_size=8)
    env = B200VecEnv(k=10, frag_len=L, max_new_nodes=48, max_improvement_step=0)
    p = _gpu_policy(sd)
    s = env.reset(x)
    steps = 0
    while not env.done:
        a = p(s)
        s, r, d = env.step(a, hook)
        steps += 1
        assert steps < 200
    for e in range(E):
        t = env.state.current_tour(e)
        assert sorted(t) == list(range(N))
        xt = x[e, t].double().cpu()
        ln = float((xt - xt.roll(-1, 0)).pow(2).sum(1).sqrt().sum())
        assert abs(ln - float(env.state.cur_len[e])) <= 1e-4 * ln
