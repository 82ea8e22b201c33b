"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference at /root/reference.

The reference (Learning4Optimization-HUST/H-TSP) is pure Python and imports a number of
packages that are absent from this image but are not used by the lower-level path-solver
hot path (pytorch_lightning, hydra, omegaconf, dgl, torch_geometric, torch_scatter,
matplotlib, tsplib95, lkh, ...).  This module injects inert ``sys.modules`` stand-ins for
them and then imports the reference *as is* so that its own code produces golden vectors
(`oracle/make_golden.py`) and pins the restatement in `oracle/path_solver_oracle.py`.

/root/reference only exists in the build container, never on the GPU box: nothing in the
`-m gpu` tests, `smoke()` or `bench.py` imports this file.  `available()` says whether the
reference tree is mounted.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("HTSP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "rl4cop", "train_path_solver.py"))


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__dict__["__path__"] = []  # behave like a package so "import a.b" works
        sys.modules[name] = mod
    for key, val in attrs.items():
        setattr(mod, key, val)
    return mod


def _install_stubs() -> None:
    import torch

    class LightningModule(torch.nn.Module):
        """Just enough of pl.LightningModule for PathSolver / HTSP_PPO construction."""

        def save_hyperparameters(self, *a, **k):
            return None

        def log(self, *a, **k):
            return None

        def log_dict(self, *a, **k):
            return None

        def print(self, *a, **k):
            return None

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    def _have(name: str) -> bool:
        try:
            importlib.import_module(name)
            return True
        except Exception:
            return False

    if not _have("pytorch_lightning"):
        _stub("pytorch_lightning", LightningModule=LightningModule, Trainer=_Dummy,
              seed_everything=lambda *a, **k: None)
        _stub("pytorch_lightning.loggers", WandbLogger=_Dummy, TensorBoardLogger=_Dummy)
        _stub("pytorch_lightning.callbacks", ModelCheckpoint=_Dummy, LearningRateMonitor=_Dummy)
        _stub("pytorch_lightning.plugins", DDPPlugin=_Dummy)
    if not _have("hydra"):
        def _main(*a, **k):
            def deco(fn):
                return fn
            return deco
        _stub("hydra", main=_main)
    if not _have("omegaconf"):
        _stub("omegaconf", DictConfig=dict, OmegaConf=_Dummy)
    for name in ("dgl", "dgl.function", "torch_geometric", "torch_geometric.nn",
                 "tsplib95", "lkh", "fast_pytorch_kmeans", "mistree"):
        if not _have(name):
            _stub(name)
    if not _have("torch_scatter"):
        # pytorch-scatter 2.0.9 (requirements.txt:159) is absent.  The upper policy calls exactly two of its
        # functions (rl_models.py:217,230), both with dim=0 and a dense 0..G-1 index from torch.unique; their
        # published semantics, restated with stock torch ops:
        #   scatter_mean(src, index, dim=0): out[g] = sum(src[index == g]) / max(count(index == g), 1)
        #   scatter_max(src, index, dim=0):  (out, argmax), out[g] = max(src[index == g])
        def scatter_mean(src, index, dim=0):
            assert dim == 0 and index.ndim == 1
            n = int(index.max()) + 1
            out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add_(0, index, src)
            cnt = torch.zeros(n, dtype=src.dtype, device=src.device).index_add_(
                0, index, torch.ones(index.shape[0], dtype=src.dtype, device=src.device))
            return out / cnt.clamp(min=1).view((n,) + (1,) * (src.ndim - 1))

        def scatter_max(src, index, dim=0):
            assert dim == 0 and index.ndim == 1 and src.ndim == 2
            n = int(index.max()) + 1
            out = torch.zeros(n, src.shape[1], dtype=src.dtype, device=src.device)
            out.scatter_reduce_(0, index[:, None].expand_as(src), src, "amax", include_self=False)
            return out, None

        _stub("torch_scatter", scatter_mean=scatter_mean, scatter_max=scatter_max)
    if not _have("matplotlib"):
        _stub("matplotlib", use=lambda *a, **k: None)
        _stub("matplotlib.pyplot")
        _stub("matplotlib.collections", LineCollection=_Dummy)
    if not _have("wandb"):
        _stub("wandb", Image=_Dummy)
    if not _have("networkx"):
        _stub("networkx")


_loaded = {}


def load_path_solver_module():
    """Returns the reference module ``rl4cop.train_path_solver`` (imported unmodified)."""
    if "tps" in _loaded:
        return _loaded["tps"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    tps = importlib.import_module("rl4cop.train_path_solver")
    _loaded["tps"] = tps
    return tps


def load_h_tsp_module():
    """Returns the reference module ``h_tsp`` (RLSolver hook, Env, VecEnv, LargeState)."""
    if "h_tsp" in _loaded:
        return _loaded["h_tsp"]
    load_path_solver_module()
    h_tsp = importlib.import_module("h_tsp")
    _loaded["h_tsp"] = h_tsp
    return h_tsp


def load_rl_models_module():
    """Returns the reference module ``rl_models`` (IMPALAEncoder, ActorPPO: the upper-level policy)."""
    if "rl_models" in _loaded:
        return _loaded["rl_models"]
    load_path_solver_module()
    _loaded["rl_models"] = importlib.import_module("rl_models")
    return _loaded["rl_models"]


def make_cfg(n_layers: int = 6, graph_size: int = 200, group_size: int = 200,
             embedding_dim: int = 128, n_heads: int = 8) -> SimpleNamespace:
    """The fields PathSolver.__init__/forward read (rl4cop/train_path_solver.py:171-211,234,245)."""
    return SimpleNamespace(
        node_dim=2, val_type="x8Aug_nTraj", encoder_type="mha", n_layers=n_layers,
        n_heads=n_heads, embedding_dim=embedding_dim, add_init_projection=True,
        tanh_clipping=10, n_decoding_neighbors=None, graph_size=graph_size,
        group_size=group_size, precision=32, learning_rate=1e-3, weight_decay=1e-6,
        gpus=[0], random_range=0, gnn_framework="pyg",
    )


def build_reference_path_solver(state_dict=None, **cfg_kw):
    """Constructs the reference PathSolver (eval mode, fp32, CPU) and optionally loads weights."""
    import torch
    tps = load_path_solver_module()
    model = tps.PathSolver(make_cfg(**cfg_kw))
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        # num_batches_tracked is the only key family we do not generate
        assert all(k.endswith("num_batches_tracked") for k in missing), missing
        assert not unexpected, unexpected
    model.eval()
    torch.backends.cuda.matmul.allow_tf32 = False
    return model


class SoftmaxTap:
    """Records the input of every F.softmax call made by the reference decoder
    (rl4cop/models/models.py:442 -> the masked, tanh-clipped logits of each decode step)
    without editing the reference: swaps the module-level name ``F`` of ``models.models``
    for a forwarding proxy while active."""

    def __init__(self):
        self.logits = []

    def __enter__(self):
        import torch.nn.functional as realF
        tap = self

        class _Proxy:
            def __getattr__(self, name):
                return getattr(realF, name)

            def softmax(self, inp, *a, **k):
                tap.logits.append(inp.detach().clone())
                return realF.softmax(inp, *a, **k)

        self._mod = sys.modules["models.models"]
        self._saved = self._mod.F
        self._mod.F = _Proxy()
        return self

    def __exit__(self, *exc):
        self._mod.F = self._saved
        return False
