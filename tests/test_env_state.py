"""Environment scoring step (Env._step / LargeState.move_to /
get_tour_distance / update_neighbor_coord_ / mask, h_tsp.py:288-340,569-603, rl_utils.py:87-159).

CPU tests: the oracle restatement against vectors recorded from the unmodified reference on real
VecEnv episodes (tests/golden/env_state_golden.npz).  GPU tests (-m gpu): the device-resident state of
h-tsp_b200/env_state.py (C ABI kernels) against the oracle and the golden vectors -- tours, masks and
neighbour coordinates bit-exact, tour lengths / rewards within 1e-5 relative of the tour length."""
import os

import numpy as np
import pytest
import torch

from oracle import env_state_oracle as eo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def egold():
    return np.load(os.path.join(ROOT, "tests", "golden", "env_state_golden.npz"))


def _replay(make_state, step, egold, tag, check_state):
    x, knn = egold[f"{tag}_x"], egold[f"{tag}_knn"]
    st = make_state(x, knn, egold[f"{tag}_init_tour"].tolist())
    check_state(st, egold[f"{tag}_init_state"], egold[f"{tag}_init_tour"], float(egold[f"{tag}_init_len"]))
    keep = {int(s): i for i, s in enumerate(egold[f"{tag}_state_steps"])}
    n_wrap = 0
    for i in range(egold[f"{tag}_paths"].shape[0]):
        p = egold[f"{tag}_paths"][i, :egold[f"{tag}_path_len"][i]].tolist()
        want_tour = egold[f"{tag}_tours"][i, :egold[f"{tag}_tour_len"][i]]
        _, reward = step(st, p)
        length = float(egold[f"{tag}_len"][i])
        check_state(st, egold[f"{tag}_states"][keep[i]] if i in keep else None, want_tour, length)
        # reward = old_len - new_len: a difference of two fp32 tour lengths, each good to 1e-5 relative
        assert abs(float(reward) - float(egold[f"{tag}_reward"][i])) <= 2e-5 * length
        n_wrap += int(want_tour[-1] == p[-1] and want_tour[0] != p[0] and len(want_tour) > len(p))
    return n_wrap


@pytest.mark.parametrize("tag", ["n300", "n1000"])
def test_oracle_matches_reference_golden(egold, tag):
    def check(st, state, tour, length):
        assert st.current_tour == [int(c) for c in tour]
        assert abs(float(st.current_tour_len) - length) <= 1e-6 * max(length, 1.0)
        if state is not None:
            assert np.array_equal(st.to_tensor(), state)

    _replay(lambda x, knn, t: eo.LargeStateOracle(x, knn, t), eo.env_step, egold, tag, check)


def test_oracle_matches_live_reference_when_mounted():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference is not mounted (GPU box)")
    H = ref_loader.load_h_tsp_module()
    g = torch.Generator().manual_seed(3)
    x = torch.rand(120, 2, generator=g)
    env = H.Env(k=6, frag_len=30, max_new_nodes=20, max_improvement_step=0)
    env.reset(x)
    st = eo.LargeStateOracle(x.numpy(), env.state.knn_neighbor.numpy(), list(env.state.current_tour))
    assert np.array_equal(st.to_tensor(), env.state.to_tensor().numpy())
    rng = np.random.default_rng(0)
    for _ in range(6):      # splice random windows of the tour, re-ordered, plus unseen cities
        tour = list(env.state.current_tour)
        unseen = [c for c in range(120) if c not in tour]
        i, w = int(rng.integers(0, len(tour))), min(len(tour), 6)
        window = [tour[(i + j) % len(tour)] for j in range(w)]
        inner = window[1:-1] + [int(c) for c in rng.permutation(unseen)[:10]]
        path = [window[0]] + [int(c) for c in rng.permutation(inner)] + [window[-1]]
        s_ref, r_ref, _ = env._step(path)
        s_or, r_or = eo.env_step(st, path)
        assert st.current_tour == env.state.current_tour
        assert np.array_equal(s_or, s_ref.numpy())
        assert abs(float(r_or) - float(r_ref)) <= 2e-5 * float(env.state.current_tour_len)


# ------------------------------------------------------------------------------------ GPU
def _gpu_check_factory():
    def check(st, state, tour, length):
        assert st.current_tour == [int(c) for c in tour]                 # integer work: bit-exact
        assert st.current_num_cities == len(tour)
        # fp32 sum of fp32-rounded fp64 edge lengths in another summation order: 1e-5 relative (north_star)
        assert abs(float(st.current_tour_len) - length) <= 1e-5 * max(length, 1.0)
        if state is not None:
            got = st.to_tensor().cpu().numpy()
            assert np.array_equal(got, state)                             # coords, masks, neighbour rows: bit-exact
    return check


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["n300", "n1000"])
def test_b200_state_matches_reference_golden(built_lib, egold, tag):
    from htsp_b200.env_state import B200LargeState

    def make(x, knn, t):
        return B200LargeState(torch.from_numpy(x).cuda(), knn.shape[1], t, torch.from_numpy(knn).cuda(),
                              max_path_len=int(egold[f"{tag}_paths"].shape[1]))

    def step(st, p):
        r = st.move_to(p)
        return st.to_tensor(), float(r)

    _replay(make, step, egold, tag, _gpu_check_factory())


def _random_episode(N, k, frag, rng):
    """Random instance + a valid sequence of splice paths (windows of the tour re-ordered with unseen
    cities mixed in, both splice branches), produced with the oracle."""
    x = rng.random((N, 2), dtype=np.float32)
    d = ((x[:, None, :].astype(np.float64) - x[None].astype(np.float64)) ** 2).sum(-1)
    np.fill_diagonal(d, np.inf)
    knn = np.argsort(d, axis=1, kind="stable")[:, :k]
    init = [0, int(knn[0, 0])]
    st = eo.LargeStateOracle(x, knn, init)
    paths = []
    while st.current_num_cities < N:
        tour = st.current_tour
        unseen = np.setdiff1d(np.arange(N), tour)
        i = int(rng.integers(0, len(tour)))
        w = min(len(tour), max(2, frag // 4))
        window = [tour[(i + j) % len(tour)] for j in range(w)]
        new = [int(c) for c in rng.permutation(unseen)[:frag - w]]
        inner = [int(c) for c in rng.permutation(window[1:-1] + new)]
        p = [window[0]] + inner + [window[-1]]
        eo.env_step(st, p)
        paths.append(p)
    return x, knn, init, paths


@pytest.mark.gpu
def test_b200_vec_state_matches_oracle_batched(built_lib):
    """E = 3 different instances stepped together (one launch pair per step) against 3 oracles;
    environments that finish early keep receiving an improvement splice of an existing window."""
    from htsp_b200.env_state import B200VecState
    rng = np.random.default_rng(5)
    eps = [_random_episode(N, 7, 40, rng) for N in (400, 400, 400)]
    xs = torch.from_numpy(np.stack([e[0] for e in eps])).cuda()
    knn = torch.from_numpy(np.stack([e[1] for e in eps])).cuda()
    v = B200VecState(xs, knn, [e[2] for e in eps], max_path_len=40)
    oracles = [eo.LargeStateOracle(e[0], e[1], e[2]) for e in eps]
    steps = max(len(e[3]) for e in eps)
    n_wrap = 0
    for s in range(steps):
        paths = []
        for e, o in zip(eps, oracles):
            if s < len(e[3]):
                paths.append(e[3][s])
            else:               # improvement step on a finished tour: re-order an 8-city window
                t = o.current_tour
                i = int(rng.integers(0, len(t)))
                win = [t[(i + j) % len(t)] for j in range(8)]
                paths.append([win[0]] + [int(c) for c in rng.permutation(win[1:-1])] + [win[-1]])
        olds = [list(o.current_tour) for o in oracles]
        want = [eo.env_step(o, p) for o, p in zip(oracles, paths)]
        state, reward = v.step(paths)
        state, reward = state.cpu().numpy(), reward.cpu().numpy()
        for e in range(3):
            assert v.current_tour(e) == oracles[e].current_tour
            assert np.array_equal(state[e], want[e][0])
            assert abs(float(reward[e]) - float(want[e][1])) <= 2e-5 * float(oracles[e].current_tour_len)
            n_wrap += int(olds[e].index(paths[e][-1]) < olds[e].index(paths[e][0]))
    assert n_wrap > 0, "the wrap-around splice branch (h_tsp.py:304-307) must be exercised"


@pytest.mark.gpu
def test_b200_state_asserts_like_the_reference(built_lib):
    from htsp_b200.env_state import B200LargeState
    rng = np.random.default_rng(1)
    x, knn, init, paths = _random_episode(60, 5, 12, rng)

    def fresh():
        st = B200LargeState(torch.from_numpy(x).cuda(), 5, init, torch.from_numpy(knn).cuda(), max_path_len=12)
        st.move_to(paths[0])
        return st

    tour = fresh().current_tour
    outside = [c for c in range(60) if c not in tour]
    with pytest.raises(AssertionError, match="endpoint"):          # h_tsp.py:293-294 (list.index raises there)
        fresh().move_to([outside[0], outside[1], tour[0]])
    with pytest.raises(AssertionError, match="coincide"):          # h_tsp.py:295
        fresh().move_to([tour[1], outside[0], tour[1]])
    with pytest.raises(AssertionError, match="duplicate"):         # h_tsp.py:309-311
        fresh().move_to([tour[0], tour[3], tour[1]])


@pytest.mark.gpu
def test_failed_splice_keeps_the_old_tour(built_lib):
    """With check=False a failed splice (non-zero status) must leave that environment as it was: the double buffer is
    flipped for all environments at once, so the kernel carries the old tour over."""
    from htsp_b200.env_state import B200VecState
    rng = np.random.default_rng(2)
    x, knn, init, paths = _random_episode(60, 5, 12, rng)
    xs = torch.from_numpy(np.stack([x, x])).cuda()
    ks = torch.from_numpy(np.stack([knn, knn])).cuda()
    st = B200VecState(xs, ks, [init, init], 12)
    st.step([paths[0], paths[0]])
    before = st.current_tour(1)
    outside = [c for c in range(60) if c not in before]
    good = paths[1]
    bad = [outside[0], outside[1], before[0]]            # first endpoint is not on the tour (h_tsp.py:293-294)
    st.step([good, bad], check=False)
    assert st.status.cpu().tolist()[0] == 0 and st.status.cpu().tolist()[1] != 0
    assert st.current_tour(1) == before                  # untouched
    assert len(st.current_tour(0)) > len(before)         # the other environment advanced


@pytest.mark.gpu
def test_b200_state_tsp10000_properties(built_lib):
    """TSP-10000-sized instance, fragments of 200 (evaluate.py defaults): after every step the tour
    is duplicate-free, the city->position index is its inverse, the length equals the O(N^2)-free
    recomputation in float64 within 1e-5, and selected == membership of the tour."""
    from htsp_b200.env_state import B200VecState
    rng = np.random.default_rng(11)
    N, k = 10000, 40
    x = rng.random((N, 2), dtype=np.float32)
    xt = torch.from_numpy(x).cuda()
    # k-NN by a coarse grid search on the device (test plumbing; Env.reset does this in the reference)
    d = torch.cdist(xt.double(), xt.double())
    d.fill_diagonal_(float("inf"))
    knn = d.topk(k, largest=False).indices
    del d
    v = B200VecState(xt[None], knn[None], [[0, int(knn[0, 0])]], max_path_len=200)
    total = v.cur_len.clone()
    for s in range(12):
        tour = v.current_tour(0)
        unseen = np.setdiff1d(np.arange(N), tour)
        i = int(rng.integers(0, len(tour)))
        w = min(len(tour), 10)
        window = [tour[(i + j) % len(tour)] for j in range(w)]
        new = [int(c) for c in rng.permutation(unseen)[:200 - w]]
        p = [window[0]] + [int(c) for c in rng.permutation(window[1:-1] + new)] + [window[-1]]
        _, reward = v.step([p])
        total = total - reward
        t = np.array(v.current_tour(0))
        assert len(np.unique(t)) == len(t)
        pos = v.pos[0].cpu().numpy()
        assert np.array_equal(pos[t], np.arange(len(t))) and (pos >= 0).sum() == len(t)
        a, b = x[t].astype(np.float64) * 1000, x[np.roll(t, -1)].astype(np.float64) * 1000
        want = float(np.sqrt(((a - b) ** 2).sum(1)).sum() / 1000)
        assert abs(float(v.cur_len[0]) - want) <= 1e-5 * want
        assert abs(float(total[0]) - want) <= 1e-4 * want      # rewards telescope to the tour length
        sel = v.selected[0].cpu().numpy().astype(bool)
        assert sel.sum() == len(t) and sel[t].all()


# ------------------------------------------------------------------------------------ fragment selection (8f-2)
@pytest.mark.parametrize("tag", ["n300", "n1000"])
def test_fragment_oracle_matches_reference_golden(egold, tag):
    """Env.reset k-NN and Env.get_fragment_knn / _extend_fragment of the oracle against what the
    unmodified reference produced on the recorded episodes (every step, exact integer equality)."""
    x = egold[f"{tag}_x"]
    frag_len, max_new, k = [int(v) for v in egold[f"{tag}_cfg"]]
    knn, init = eo.knn_reset(x, k)
    assert np.array_equal(knn, egold[f"{tag}_knn"]) and init == egold[f"{tag}_init_tour"].tolist()
    st = eo.LargeStateOracle(x, knn, init)
    for i in range(egold[f"{tag}_paths"].shape[0]):
        frag, new_cities, start = eo.fragment_knn(st, egold[f"{tag}_actions"][i], frag_len, max_new)
        assert frag == egold[f"{tag}_fragments"][i].tolist()
        want_new = egold[f"{tag}_new_cities"][i]
        assert new_cities == want_new[want_new >= 0].tolist()
        assert np.array_equal(x[start] * 2 - 1, egold[f"{tag}_start_coord"][i])     # Env.unscale_action
        eo.env_step(st, egold[f"{tag}_paths"][i, :egold[f"{tag}_path_len"][i]].tolist())


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["n300", "n1000"])
def test_b200_knn_and_fragments_match_reference_golden(built_lib, egold, tag):
    """K7 against the unmodified reference's recorded episodes: the k-NN lists of Env.reset and, step
    by step, the fragment / new cities / start city of Env.get_fragment_knn -- exact integers."""
    from htsp_b200.vec_env import B200VecEnv
    x = torch.from_numpy(egold[f"{tag}_x"]).cuda()
    frag_len, max_new, k = [int(v) for v in egold[f"{tag}_cfg"]]
    env = B200VecEnv(k=k, frag_len=frag_len, max_new_nodes=max_new, max_improvement_step=2)
    env.reset(x[None])
    assert np.array_equal(env.knn[0].cpu().numpy(), egold[f"{tag}_knn"])
    assert env.state.current_tour(0) == egold[f"{tag}_init_tour"].tolist()
    for i in range(egold[f"{tag}_paths"].shape[0]):
        act = torch.from_numpy(egold[f"{tag}_actions"][i])[None].cuda()
        frag, newc, n_new, start = env.fragments(act, scaled=True)
        assert int(env._fstatus[0]) == 0
        assert frag[0].cpu().tolist() == egold[f"{tag}_fragments"][i].tolist()
        want_new = egold[f"{tag}_new_cities"][i]
        assert int(n_new[0]) == int((want_new >= 0).sum())
        assert newc[0].cpu().numpy()[: int(n_new[0])].tolist() == want_new[want_new >= 0].tolist()
        got_start = egold[f"{tag}_x"][int(start[0])] * 2 - 1
        assert np.array_equal(got_start, egold[f"{tag}_start_coord"][i])
        env.state.step([egold[f"{tag}_paths"][i, :egold[f"{tag}_path_len"][i]].tolist()])


@pytest.mark.gpu
def test_b200_vec_env_full_loop_against_oracle(built_lib, sd6):
    """The whole device-resident construction loop (fragment selection -> path solver hook -> scoring)
    for 3 environments, mirrored step by step by the oracle driven with the SAME paths: fragments,
    tours, state tensors identical; the episode ends with complete, duplicate-free tours."""
    from htsp_b200 import B200PathSolver, B200RLSolver
    from htsp_b200.vec_env import B200VecEnv
    N, k, FL, MAXNEW = 260, 8, 40, 30
    g = torch.Generator().manual_seed(21)
    x = torch.rand(3, N, 2, generator=g)
    model = B200PathSolver(state_dict=sd6, mode="x3", device="cuda", graph_size=FL)
    hook = B200RLSolver(model, sample_size=8)
    env = B200VecEnv(k=k, frag_len=FL, max_new_nodes=MAXNEW, max_improvement_step=1)
    env.reset(x.cuda())
    oracles = []
    for e in range(3):
        knn, init = eo.knn_reset(x[e].numpy(), k)
        assert np.array_equal(knn, env.knn[e].cpu().numpy())
        oracles.append(eo.LargeStateOracle(x[e].numpy(), knn, init))
    steps = 0
    while not env.done:
        act = env.random_action()
        active = (~env.dones).cpu().tolist()
        frag, _, _, _ = env.fragments(act, ~env.dones)
        scaled = (act * 0.5 + 0.5).cpu().numpy()
        for e in range(3):
            if active[e]:
                want, _, _ = eo.fragment_knn(oracles[e], scaled[e], FL, MAXNEW)
                assert frag[e].cpu().tolist() == want
        states, rewards, dones = env.step(act, hook)
        for e in range(3):
            if active[e]:
                path = env.state._paths[e].cpu().tolist()
                st, r = eo.env_step(oracles[e], path)
                assert env.state.current_tour(e) == oracles[e].current_tour
                assert np.array_equal(states[e].cpu().numpy(), st)
                assert abs(float(rewards[e]) - float(r)) <= 2e-5 * float(oracles[e].current_tour_len)
        steps += 1
        assert steps < 100
    for e in range(3):
        t = env.state.current_tour(e)
        assert sorted(t) == list(range(N))
