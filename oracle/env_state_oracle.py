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


# ------------------------------------------------------------------------------------------------
# fragment selection: Env.reset k-NN (h_tsp.py:443-469),
# rl_utils.get_nearest_city_idx (rl_utils.py:46-62), Env.get_fragment_knn (h_tsp.py:482-542) with
# the numba FIFO expansion (h_tsp.py:472-480) and Env._extend_fragment (h_tsp.py:638-661)
# ------------------------------------------------------------------------------------------------
def knn_reset(x: np.ndarray, k: int) -> Tuple[np.ndarray, List[int]]:
    """k nearest neighbours of every city under fp32(fp64 cdist x 1000) with an infinite diagonal,
    ascending (torch.topk(largest=False)); equal distances are ordered by city index here (torch leaves
    the order of exact ties unspecified).  Also the initial 2-city tour [0, knn[0][0]] (h_tsp.py:456-459)."""
    xs = x.astype(np.float64) * DISTANCE_SCALE
    n = x.shape[0]
    knn = np.empty((n, k), dtype=np.int64)
    for i in range(n):
        d = np.sqrt(((xs - xs[i]) ** 2).sum(axis=1)).astype(np.float32)
        d[i] = np.inf
        knn[i] = np.lexsort((np.arange(n), d))[:k]
    return knn, [0, int(knn[0, 0])]


def nearest_city_idx(x: np.ndarray, coord: np.ndarray, mask: np.ndarray) -> int:
    """rl_utils.get_nearest_city_idx: arg-min over the masked cities of float32(fp64 distance to
    `coord`), first (lowest) city index among equal distances."""
    cities = np.nonzero(mask)[0]
    d = np.sqrt(((x[cities].astype(np.float64) - coord.astype(np.float64)[None]) ** 2).sum(axis=1)).astype(np.float32)
    return int(cities[int(np.argmin(d))])


def extend_fragment(tour: List[int], nearest_city: int, new_cities: List[int], frag_len: int) -> List[int]:
    nearest_idx = tour.index(nearest_city)
    total = frag_len - len(new_cities)
    assert total > 0
    offset = nearest_idx - total // 2
    reorder = tour[offset:] + tour[:offset]            # Python slice semantics (negative offset wraps)
    assert len(reorder) == len(tour)
    frag = reorder[:total // 2] + list(new_cities) + reorder[total // 2:total]
    assert len(frag) == frag_len and len(set(frag)) == frag_len
    return frag


def fragment_knn(state: LargeStateOracle, coord: np.ndarray, frag_len: int, max_new_nodes: int):
    """Env.get_fragment_knn for an environment that is not done: (fragment, new_cities, start_city).
    `coord` is the scaled action in [0,1]^2 (Env.scale_action has been applied by the caller)."""
    x = state.x
    construction_done = state.current_num_cities == state.graph_size
    if not construction_done:
        nearest_new = nearest_city_idx(x, coord, ~state.selected_mask)
        if state.current_num_cities != 0:
            start_city = nearest_city_idx(x, x[nearest_new], state.selected_mask)
            new_start = nearest_city_idx(x, x[start_city], ~state.selected_mask)
        else:
            start_city, new_start = None, nearest_new
        max_cities = max(max_new_nodes, frag_len - state.current_num_cities)
        new_cities, queue, seen = [], [new_start], {new_start}
        while queue and len(new_cities) < max_cities:
            node = queue.pop(0)
            new_cities.append(int(node))
            for n in state.knn_neighbor[node]:
                n = int(n)
                if n not in seen and not state.selected_mask[n]:
                    queue.append(n)
                    seen.add(n)
    else:
        start_city = nearest_city_idx(x, coord, state.selected_mask)
        new_cities = []
    if state.current_num_cities != 0:
        fragment = extend_fragment(state.current_tour, start_city, new_cities, frag_len)
    else:
        fragment = list(new_cities)
    return fragment, new_cities, start_city
