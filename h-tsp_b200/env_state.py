"""Device-resident state of the large-TSP environment: the per-step
bookkeeping of the reference's `Env._step` (h_tsp.py:569-603) -- `LargeState.move_to` list splice
(:288-320), `rl_utils.get_tour_distance` x 2 over an Ntot x Ntot matrix (rl_utils.py:87-102),
`rl_utils.update_neighbor_coord_` (rl_utils.py:135-159), `LargeState.mask` (:330-340) and
`LargeState.to_tensor` (:396-408) -- as two kernel launches for ALL environments of a VecEnv
(`htsp_env_splice`, `htsp_env_update`, include/htsp_b200.h), O(T) per environment and without the
distance matrix (1.2 GB per TSP-10000 instance in the reference).

`B200VecState` is the batched object; `B200LargeState` is a one-environment view with the
reference's attribute names (`current_tour`, `current_tour_len`, `current_num_cities`,
`selected_mask`, `available_mask`, `neighbor_coord`, `move_to`, `to_tensor`) so that the reference's
fragment selection (`Env.get_fragment_knn`, still the reference's own code: section 8f rank 2) keeps
reading it.  torch provides device memory and streams only; there is no CPU path."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _abi

_STATUS = {1: "an endpoint of the new path is not on the current tour (h_tsp.py:293-294)",
           2: "start and end of the new path coincide (h_tsp.py:295)",
           3: "duplicate city in the tour (h_tsp.py:309-311)",
           4: "the tour did not grow (h_tsp.py:313-316)",
           5: "city index out of range"}


class B200VecState:
    """State of E environments over instances x [E,Ntot,2] with k-NN lists knn [E,Ntot,k]
    (`Env.reset`, h_tsp.py:443-469, computes them; kept as an input here)."""

    def __init__(self, x: torch.Tensor, knn: torch.Tensor, init_tours: Sequence[Sequence[int]],
                 max_path_len: int = 200):
        assert x.is_cuda, "htsp_b200 has no CPU path"
        self.device = x.device
        self.x = x.to(torch.float32).contiguous()
        E, N, C = self.x.shape
        assert C == 2 and knn.shape[:2] == (E, N) and len(init_tours) == E
        self.E, self.N, self.k, self.P = E, N, int(knn.shape[2]), int(max_path_len)
        self.knn = knn.to(self.device, torch.int64).contiguous()
        dev = self.device
        self._tours = [torch.full((E, N), -1, dtype=torch.int64, device=dev) for _ in range(2)]
        self._cur = 0
        self.tour_len = torch.zeros(E, dtype=torch.int32, device=dev)
        self.pos = torch.full((E, N), -1, dtype=torch.int32, device=dev)
        self.selected = torch.zeros(E, N, dtype=torch.uint8, device=dev)
        self.available = torch.zeros(E, N, dtype=torch.uint8, device=dev)
        self.neighbor_coord = torch.zeros(E, N, 4, dtype=torch.float32, device=dev)
        self.state = torch.empty(E, N, 8, dtype=torch.float32, device=dev)
        self.cur_len = torch.zeros(E, dtype=torch.float32, device=dev)
        self.reward = torch.zeros(E, dtype=torch.float32, device=dev)
        self.status = torch.zeros(E, dtype=torch.int32, device=dev)
        self._paths = torch.zeros(E, self.P, dtype=torch.int64, device=dev)
        self._path_len = torch.zeros(E, dtype=torch.int32, device=dev)
        # initial 2-city (or empty) tours: LargeState.__init__ (h_tsp.py:278-286)
        for e, t in enumerate(init_tours):
            assert len(t) in (0, 2)
            if len(t):
                tt = torch.tensor(list(t), dtype=torch.int64, device=dev)
                self._tours[0][e, :len(t)] = tt
                self.pos[e, tt] = torch.arange(len(t), dtype=torch.int32, device=dev)
                self._paths[e, :len(t)] = tt
            self.tour_len[e] = len(t)
            self._path_len[e] = len(t)
        self._update()
        self.reward.zero_()

    # -- kernels ----------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _update(self) -> None:
        lib = _abi.lib()
        with torch.cuda.device(self.device):
            _abi.check(lib.htsp_env_update(self.x.data_ptr(), self.tours.data_ptr(), self.tour_len.data_ptr(),
                                           self._paths.data_ptr(), self._path_len.data_ptr(),
                                           self.knn.data_ptr(), self.k, self.E, self.P, self.N,
                                           self.selected.data_ptr(), self.available.data_ptr(),
                                           self.neighbor_coord.data_ptr(), self.state.data_ptr(),
                                           self.cur_len.data_ptr(), self.reward.data_ptr(), self._stream()),
                       "htsp_env_update")

    @property
    def tours(self) -> torch.Tensor:
        """[E,Ntot] i64, the first tour_len[e] entries of row e are the current tour."""
        return self._tours[self._cur]

    def set_paths(self, new_paths: Sequence[Sequence[int]]) -> None:
        """Host lists -> the device path buffer (one pinned-less H2D copy of E*P ids)."""
        assert len(new_paths) == self.E
        lens = [len(p) for p in new_paths]
        assert max(lens) <= self.P, "path longer than max_path_len"
        buf = torch.zeros(self.E, self.P, dtype=torch.int64)
        for e, p in enumerate(new_paths):
            buf[e, :lens[e]] = torch.as_tensor(list(p), dtype=torch.int64)
        self._paths.copy_(buf, non_blocking=False)
        self._path_len.copy_(torch.tensor(lens, dtype=torch.int32))

    def step_device(self, check: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """One `Env._step` for every environment on the paths currently in the device buffer
        (`_paths` [E,P] i64, `_path_len` [E] i32): -> (state [E,Ntot,8] f32, reward [E] f32)."""
        lib = _abi.lib()
        nxt = 1 - self._cur
        with torch.cuda.device(self.device):
            _abi.check(lib.htsp_env_splice(self._paths.data_ptr(), self._path_len.data_ptr(), self.E, self.P,
                                           self.N, self._tours[self._cur].data_ptr(),
                                           self._tours[nxt].data_ptr(), self.tour_len.data_ptr(),
                                           self.pos.data_ptr(), self.status.data_ptr(), self._stream()),
                       "htsp_env_splice")
        if check:      # the reference asserts on the host (h_tsp.py:293-320); one 4*E byte D2H
            st = self.status.cpu().tolist()
            for e, s in enumerate(st):
                assert s == 0, f"environment {e}: {_STATUS.get(s, s)}"
        self._cur = nxt
        self._update()
        return self.state, self.reward

    def step(self, new_paths: Sequence[Sequence[int]], check: bool = True):
        self.set_paths(new_paths)
        return self.step_device(check)

    def current_tour(self, e: int) -> List[int]:
        n = int(self.tour_len[e])
        return self.tours[e, :n].cpu().tolist()


class B200LargeState:
    """One environment with the attribute names of the reference's `LargeState` (h_tsp.py:231-408)."""

    def __init__(self, x: torch.Tensor, k: int, init_tour: List[int], knn_neighbor: torch.Tensor,
                 max_path_len: int = 200):
        assert x.ndim == 2
        self.x = x
        self.device = x.device
        self.graph_size = x.shape[0]
        self.k = k
        self.knn_neighbor = knn_neighbor
        self._v = B200VecState(x[None], knn_neighbor[None], [init_tour], max_path_len)
        self._tour_cache: Optional[List[int]] = None

    @property
    def current_tour(self) -> List[int]:
        if self._tour_cache is None:
            self._tour_cache = self._v.current_tour(0)
        return self._tour_cache

    @property
    def current_num_cities(self) -> int:
        return int(self._v.tour_len[0])

    @property
    def current_tour_len(self) -> torch.Tensor:
        return self._v.cur_len[0]

    @property
    def selected_mask(self) -> torch.Tensor:
        return self._v.selected[0].bool()

    @property
    def available_mask(self) -> torch.Tensor:
        return self._v.available[0].bool()

    @property
    def neighbor_coord(self) -> torch.Tensor:
        return self._v.neighbor_coord[0]

    def move_to(self, new_path: List[int]) -> torch.Tensor:
        """`LargeState.move_to` + the length / reward lines of `Env._step`; returns the reward."""
        self._v.step([new_path])
        self._tour_cache = None
        return self._v.reward[0]

    def to_tensor(self) -> torch.Tensor:
        return self._v.state[0]
