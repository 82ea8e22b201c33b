"""B200VecEnv: the construction loop of the reference's `VecEnv` (h_tsp.py:684-843) with every
per-step piece on the device:

  reset(x)            Env.reset k-NN lists + 2-city tours         -> htsp_env_knn, B200VecState
  fragments(actions)  Env.get_fragment_knn for all environments   -> htsp_env_fragment
  step(actions, hook) fragments -> B200RLSolver.solve_device (the path-solver hot path) ->
                      Env._step bookkeeping (htsp_env_splice / htsp_env_update)

Nothing goes through host lists: fragment ids, solved paths, tours, masks and the [E,Ntot,8] state
tensor stay in HBM; the host reads back 4*E bytes of status per step (the reference's asserts) and the
tour lengths when asked.  The upper-level policy that produces `actions` is out of scope;
`random_action()` is the reference's own stand-in (h_tsp.py:829-830, evaluate.py:88-97)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _abi
from .env_state import B200VecState


class B200VecEnv:
    def __init__(self, k: int, frag_len: int = 200, max_new_nodes: int = 160, max_improvement_step: int = 5):
        self.k, self.frag_len = k, frag_len
        self.max_new_nodes, self.max_improvement_step = max_new_nodes, max_improvement_step

    # -- Env.reset for every instance (h_tsp.py:443-469, 704-720) -----------------------------------
    @torch.no_grad()
    def reset(self, x: torch.Tensor) -> torch.Tensor:
        assert x.is_cuda and x.ndim == 3 and x.shape[2] == 2, "htsp_b200 has no CPU path"
        lib = _abi.lib()
        self.x = x.to(torch.float32).contiguous()
        self.device = x.device
        E, N, _ = self.x.shape
        self.E, self.N = E, N
        self.knn = torch.empty(E, N, self.k, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _abi.check(lib.htsp_env_knn(self.x.data_ptr(), E, N, self.k, self.knn.data_ptr(), self._stream()),
                       "htsp_env_knn")
        first_nn = self.knn[:, 0, 0].cpu().tolist()            # init_tour = [depot, nearest neighbour] (:456-459)
        self.state = B200VecState(self.x, self.knn, [[0, int(n)] for n in first_nn], self.frag_len)
        self.improvement_steps = torch.zeros(E, dtype=torch.int32, device=self.device)
        dev = self.device
        self._fragment = torch.empty(E, self.frag_len, dtype=torch.int64, device=dev)
        self._new_cities = torch.empty(E, self.frag_len, dtype=torch.int64, device=dev)
        self._n_new = torch.empty(E, dtype=torch.int32, device=dev)
        self._start_city = torch.empty(E, dtype=torch.int32, device=dev)
        self._fstatus = torch.empty(E, dtype=torch.int32, device=dev)
        return self.state.state

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- done flags (h_tsp.py:605-615) ----------------------------------------------------------------
    @property
    def construction_done(self) -> torch.Tensor:
        return self.state.tour_len == self.N

    @property
    def dones(self) -> torch.Tensor:
        return self.construction_done & (self.improvement_steps >= self.max_improvement_step)

    @property
    def done(self) -> bool:
        return bool(self.dones.all())

    def random_action(self) -> torch.Tensor:
        return torch.rand(self.E, 2, device=self.device) * 2 - 1          # Env.unscale_action(rand)

    # -- Env.get_fragment_knn for all active environments (h_tsp.py:482-542) --------------------------
    @torch.no_grad()
    def fragments(self, actions: torch.Tensor, active: Optional[torch.Tensor] = None, scaled: bool = False):
        """actions [E,2] in [-1,1]^2 (policy output) -> (fragment [E,frag_len] i64, new_cities [E,frag_len]
        i64 (-1 padded), n_new [E] i32, start_city [E] i32), all on the device."""
        lib = _abi.lib()
        a = actions.to(self.device, torch.float32)
        a = (a if scaled else a * 0.5 + 0.5).contiguous()                             # Env.scale_action
        act = None if active is None else active.to(torch.uint8).contiguous()
        s = self.state
        with torch.cuda.device(self.device):
            _abi.check(lib.htsp_env_fragment(self.x.data_ptr(), s.selected.data_ptr(), self.knn.data_ptr(), self.k,
                                             s.tours.data_ptr(), s.tour_len.data_ptr(), s.pos.data_ptr(),
                                             a.data_ptr(), _abi.ptr(act), self.E, self.N, self.frag_len,
                                             self.max_new_nodes, self._fragment.data_ptr(),
                                             self._new_cities.data_ptr(), self._n_new.data_ptr(),
                                             self._start_city.data_ptr(), self._fstatus.data_ptr(),
                                             self._stream()), "htsp_env_fragment")
        return self._fragment, self._new_cities, self._n_new, self._start_city

    # -- VecEnv.step (h_tsp.py:724-827) without auto-reset ----------------------------------------------
    @torch.no_grad()
    def step(self, actions: torch.Tensor, hook, check: bool = True, shard_solver: bool = False, frag_buffer=None
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """One environment step for every environment that is not done: fragment selection, the
        lower-level path solver (`hook.solve_device`, B200RLSolver) and the scoring step.
        Returns (states [E,Ntot,8], rewards [E], dones [E]).  `frag_buffer` is handed to the hook, which calls its
        `update_buffer` with the normalised fragments as the reference does (h_tsp.py:192).  Deviations from
        `VecEnv.step` (h_tsp.py:783-793), documented in INTEGRATION.md: no auto-reset, no `info` dict (the fragment /
        new-city / start-city tensors of the step are `self._fragment`, `self._new_cities`, `self._start_city`), and
        `dones` stays True for environments that finished earlier (the reference reports False for them).
        shard_solver (under torch.distributed, every rank holding the SAME environments and actions): the
        subproblems of the step are solved by `sharding.solve_sharded` -- each rank rolls out its block, one NCCL
        all-gather of the paths -- and every rank applies the identical scoring step."""
        active = ~self.dones
        idx = torch.nonzero(active).squeeze(1)
        s = self.state
        if idx.numel() == 0:
            raise RuntimeError("Environment has terminated!")           # h_tsp.py:735-736
        frag, _, _, _ = self.fragments(actions, active)
        if check:
            fs = self._fstatus.cpu()
            assert bool((fs == 0).all()), f"fragment length asserts failed: {fs.tolist()}"   # h_tsp.py:641,656-658
        if shard_solver:
            from . import sharding
            paths, _ = sharding.solve_sharded(hook, self.x.index_select(0, idx), frag.index_select(0, idx))
        else:
            paths, _, status, _ = hook.solve_device(self.x.index_select(0, idx), frag.index_select(0, idx), frag_buffer)
            if check:
                assert bool((status == 0).all()), "invalid path(s) returned by the rollout"
        was_done = self.construction_done.clone()
        s._paths.zero_()
        s._path_len.zero_()
        s._paths.index_copy_(0, idx, paths)
        s._path_len.index_fill_(0, idx, self.frag_len)
        states, rewards = s.step_device(check)
        self.improvement_steps += (was_done & active).to(torch.int32)    # h_tsp.py:580-581
        return states, rewards * active.to(rewards.dtype), self.dones
