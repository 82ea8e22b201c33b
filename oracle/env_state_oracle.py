"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's environment scoring step
(SURVEY.md section 8f rank 1): `LargeState.move_to` / `mask` (h_tsp.py:288-340),
`rl_utils.get_tour_distance` (rl_utils.py:87-102) over the fp32(fp64 cdist x 1000) matrix
(h_tsp.py:449-453), `rl_utils.update_neighbor_coord_` (rl_utils.py:135-159), `LargeState.to_tensor`
(h_tsp.py:396-408) and the reward of `Env._step` (h_tsp.py:569-603).

Pinned by tests/test_env_state.py against vectors recorded from the UNMODIFIED reference running
real VecEnv episodes (oracle/make_golden_env.py -> tests/golden/env_state_golden.npz) and, where
/root/reference is mounted, against the live reference objects.  Only tests/, smoke() and bench.py
may import this module; the product path (h-tsp_b200/env_state.py) never does."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

DISTANCE_SCALE = 1000   # h_tsp.py:31


def tour_distance(x: np.ndarray, tour: List[int]) -> np.float32:
    """get_tour_distance(tour, dist_matrix) with dist_matrix = float32(cdist(float64(x)*1000)):
    fp32 sum over the closed tour of the fp32-rounded fp64 edge lengths (rl_utils.py:87-102).
    Returned in the matrix' units (NOT divided by DISTANCE_SCALE)."""
    if len(tour) == 0:
        return np.float32(0.0)
    t = np.asarray(tour, dtype=np.int64)
    a = x[t].astype(np.float64) * DISTANCE_SCALE
    b = x[np.roll(t, -1)].astype(np.float64) * DISTANCE_SCALE
    d = np.sqrt(((a - b) ** 2).sum(axis=1)).astype(np.float32)
    return np.float32(d.sum(dtype=np.float32))


class LargeStateOracle:
    """State of one large-TSP environment (h_tsp.py:231-408) with the same attribute names."""

    def __init__(self, x: np.ndarray, knn_neighbor: np.ndarray, init_tour: List[int]):
        assert x.ndim == 2 and x.shape[1] == 2
        self.x = x.astype(np.float32)
        self.graph_size = x.shape[0]
        self.knn_neighbor = np.asarray(knn_neighbor, dtype=np.int64)
        assert len(init_tour) in (0, 2)
        self.selected_mask = np.zeros(self.graph_size, dtype=bool)
        self.available_mask = np.zeros(self.graph_size, dtype=bool)
        self.neighbor_coord = np.zeros((self.graph_size, 4), dtype=np.float32)
        self.current_tour = list(init_tour)
        self.current_num_cities = len(self.current_tour)
        # h_tsp.py:283-285
        self.current_tour_len = np.float32(tour_distance(self.x, self.current_tour) / np.float32(DISTANCE_SCALE))
        self._update_neighbor_coord()
        self.mask(self.current_tour)

    # rl_utils.py:135-159: row `tour[i]` <- [x[tour[i+1]], x[tour[i-1]]]; other rows keep their values
    def _update_neighbor_coord(self) -> None:
        if not self.current_tour:
            return
        t = np.asarray(self.current_tour, dtype=np.int64)
        self.neighbor_coord[t] = np.concatenate([self.x[np.roll(t, -1)], self.x[np.roll(t, 1)]], axis=1)

    # h_tsp.py:330-340
    def mask(self, new_path: List[int]) -> None:
        p = np.asarray(new_path, dtype=np.int64)
        self.selected_mask[p] = True
        self.available_mask[p] = True
        avail = np.nonzero(self.available_mask)[0]
        self.available_mask[avail] = ~self.selected_mask[self.knn_neighbor[avail]].all(axis=1)

    # h_tsp.py:288-328
    def move_to(self, new_path: List[int]) -> None:
        new_path = [int(c) for c in new_path]
        if self.current_num_cities == 0:
            self.current_tour = list(new_path)
        else:
            start_idx = self.current_tour.index(new_path[0])
            end_idx = self.current_tour.index(new_path[-1])
            assert start_idx != end_idx, new_path
            if end_idx > start_idx:
                self.current_tour = (self.current_tour[:start_idx] + new_path
                                     + self.current_tour[end_idx + 1:])
            else:
                self.current_tour = self.current_tour[end_idx + 1:start_idx] + new_path
        assert len(set(self.current_tour)) == len(self.current_tour), "duplicate city in the tour"
        if len(self.current_tour) < self.graph_size:
            assert len(self.current_tour) > self.current_num_cities
        else:
            assert len(self.current_tour) == self.graph_size
        self.current_num_cities = len(self.current_tour)
        self.current_tour_len = tour_distance(self.x, self.current_tour)       # matrix units (:322-324)
        self._update_neighbor_coord()
        self.mask(new_path)

    # h_tsp.py:396-408
    def to_tensor(self) -> np.ndarray:
        return np.concatenate([self.x, self.selected_mask[:, None].astype(np.float32),
                               self.available_mask[:, None].astype(np.float32), self.neighbor_coord], axis=1)


def env_step(state: LargeStateOracle, new_path: List[int]) -> Tuple[np.ndarray, np.float32]:
    """Env._step with greedy_reward = average_reward = False (h_tsp.py:569-603):
    (state tensor [N,8], reward = old_len - new_len); new_len = tour distance / 1000."""
    old_len = state.current_tour_len
    state.move_to(new_path)
    state.current_tour_len = np.float32(tour_distance(state.x, state.current_tour) / np.float32(DISTANCE_SCALE))
    return state.to_tensor(), np.float32(old_len - state.current_tour_len)
