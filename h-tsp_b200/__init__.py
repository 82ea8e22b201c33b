"""htsp_b200 -- B200-native (sm_100a) implementation of the H-TSP lower-level path-solver
hot path.  See DESIGN.md."""
__version__ = "0.1.0"

from .weights import expected_keys, synthetic_state_dict  # noqa: E402,F401


def __getattr__(name):  # lazy: importing the package must not require CUDA
    if name in ("B200PathSolver",):
        from .path_solver import B200PathSolver
        return B200PathSolver
    if name in ("B200RLSolver", "LowLevelSolver", "FragmentBuffer", "make_reference_subclass"):
        from . import rl_solver
        return getattr(rl_solver, name)
    raise AttributeError(name)
