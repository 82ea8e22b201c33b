"""CPU tests of the host-side logic that does not need the GPU."""
import numpy as np
import pytest
import torch

from htsp_b200.weights import expected_keys, n_layers_of, synthetic_state_dict
from oracle import ref_loader


def test_synthetic_weights_have_reference_keys_and_are_deterministic():
    sd = synthetic_state_dict(1234, 6)
    assert list(sd) == expected_keys(6) and n_layers_of(sd) == 6
    sd2 = synthetic_state_dict(1234, 6)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)
    assert sum(v.numel() for k, v in sd.items() if "running" not in k) == 1318912


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_reference_model_accepts_the_key_layout():
    sd = synthetic_state_dict(1, 2)
    m = ref_loader.build_reference_path_solver(sd, n_layers=2)
    got = {k for k in m.state_dict() if not k.endswith("num_batches_tracked")}
    assert got == set(expected_keys(2))


def test_fragment_buffer_ring_semantics():
    from htsp_b200.rl_solver import FragmentBuffer
    fb = FragmentBuffer(5, 3, 2)
    a = torch.arange(4 * 3 * 2, dtype=torch.float32).reshape(4, 3, 2)
    fb.update_buffer(a)
    assert fb.next_idx == 4 and fb.now_len == 4 and not fb.if_full
    b = a + 100
    fb.update_buffer(b[:3])          # wraps: 1 goes to the tail, 2 to the head
    assert fb.if_full and fb.now_len == 5 and fb.next_idx == 2
    assert torch.equal(fb.frag_buffer[4], b[0]) and torch.equal(fb.frag_buffer[0:2], b[1:3])
    if ref_loader.available():
        ref_loader.load_h_tsp_module()
        import rl_utils
        rb = rl_utils.FragmengBuffer(5, 3, 2)
        rb.update_buffer(a)
        rb.update_buffer(b[:3])
        assert torch.equal(rb.frag_buffer, fb.frag_buffer) and rb.next_idx == fb.next_idx


def test_product_package_does_not_import_the_oracle():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "h-tsp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_solver_refuses_without_the_cuda_library(monkeypatch, tmp_path):
    from htsp_b200 import _abi
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(_abi, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(_abi.HtspError):
        _abi.lib()


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_make_reference_subclass_is_a_reference_rlsolver():
    """the drop-in object must pass `isinstance(solver, h_tsp.RLSolver)` (VecEnv.step,
    h_tsp.py:764) against the UNMODIFIED reference class and resolve `solve` to the B200 method (no CUDA needed:
    construction and method resolution only)."""
    h_tsp = ref_loader.load_h_tsp_module()
    from htsp_b200 import rl_solver
    from htsp_b200.path_solver import B200PathSolver

    model = B200PathSolver.__new__(B200PathSolver)      # no device work: only the type is inspected
    solver = rl_solver.make_reference_subclass(h_tsp, model, sample_size=1)
    assert isinstance(solver, h_tsp.RLSolver) and isinstance(solver, h_tsp.LowLevelSolver)
    assert isinstance(solver, rl_solver.B200RLSolver)
    mro = type(solver).__mro__
    assert mro.index(rl_solver.B200RLSolver) < mro.index(h_tsp.RLSolver)
    assert type(solver).solve is rl_solver.B200RLSolver.solve
    assert type(solver).solve is not h_tsp.RLSolver.solve
    assert solver._sample_size == 2 and solver._solver_model is model      # max(sample_size, 2), h_tsp.py:168


def test_hook_reports_a_failed_rollout_by_cause():
    """B200RLSolver.solve turns what the device returned into the reference's contract (h_tsp.py:206-228) and names
    the cause of a failure: bad indices -> IndexError (torch.gather's error in the reference),
    non-finite rewards or a repeated node with good indices -> FloatingPointError naming mode='fp32', any other failed
    path check -> the reference's AssertionError.  Host logic only: solve_device is replaced by canned tensors."""
    import numpy as np
    from htsp_b200 import rl_solver
    from htsp_b200.path_solver import B200PathSolver

    model = B200PathSolver.__new__(B200PathSolver)      # no device work
    torch.nn.Module.__init__(model)
    model.mode = "fast"
    model.first_action_override = None
    hook = rl_solver.B200RLSolver(model, sample_size=4)
    B, N = 2, 5
    paths = torch.arange(N).repeat(B, 1) + 10
    x_ok = torch.rand(B, N, 2)

    def canned(lengths, status, x_new):
        hook.solve_device = lambda data, fragment, frag_buffer=None: (paths, torch.tensor(lengths), torch.tensor(status, dtype=torch.int32), x_new)
        return hook.solve(None, None, None)

    got_paths, got_lens = canned([1.5, 2.5], [0, 0], x_ok)
    assert got_paths == paths.tolist() and np.allclose(got_lens, [1.5, 2.5])
    x_nan = x_ok.clone()
    x_nan[1] = float("nan")
    with pytest.raises(IndexError, match=r"\[1\]"):
        canned([1.5, float("nan")], [0, 0], x_nan)                   # htsp_extract_normalize flagged fragment ids
    model.first_action_override = torch.tensor([0, 1, 2, N])          # a POMO start outside [0, N)
    with pytest.raises(IndexError):
        canned([float("nan"), float("nan")], [0, 0], x_ok)
    model.first_action_override = None
    with pytest.raises(FloatingPointError, match="mode='fp32'"):
        canned([float("nan"), 2.5], [0, 0], x_ok)                    # indices fine, no finite reward
    with pytest.raises(FloatingPointError, match="fp32"):
        canned([1.5, 2.5], [0, 4], x_ok)                             # ... or a tour that repeats a node
    with pytest.raises(AssertionError, match="status"):
        canned([1.5, 2.5], [2, 0], x_ok)                             # the reference's own path asserts
