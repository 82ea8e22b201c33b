"""Tour-length scoring on the device (K5): the arithmetic of rl_utils.get_tour_distance
(rl_utils.py:87-102) over dist_matrix = fp32(cdist_fp64(1000*x)) (h_tsp.py:449-453), i.e. the
reward of Env._step (h_tsp.py:586-599), computed from coordinates in O(T) instead of two
[T,Ntot] gathers of an Ntot x Ntot matrix."""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import _abi


@torch.no_grad()
def tour_lengths(coords: torch.Tensor, tours: torch.Tensor, tour_len: torch.Tensor) -> torch.Tensor:
    """coords [E,Ntot,2] f32, tours [E,Tmax] i64 (padded), tour_len [E] i32 -> [E] f32."""
    lib = _abi.lib()
    assert coords.is_cuda, "htsp_b200 has no CPU path"
    dev = coords.device
    coords = coords.to(torch.float32).contiguous()
    tours = tours.to(dev, torch.int64).contiguous()
    tour_len = tour_len.to(dev, torch.int32).contiguous()
    E, Ntot, _ = coords.shape
    out = torch.empty(E, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _abi.check(lib.htsp_tour_length(coords.data_ptr(), tours.data_ptr(), tour_len.data_ptr(), E,
                                        Ntot, tours.shape[1], out.data_ptr(),
                                        torch.cuda.current_stream(dev).cuda_stream),
                   "htsp_tour_length")
    return out


def tour_length(coords: torch.Tensor, tour: Sequence[int]) -> float:
    """Single-tour convenience with the reference's call shape (list of ints)."""
    t = torch.tensor(list(tour), dtype=torch.int64, device=coords.device)[None]
    n = torch.tensor([t.shape[1]], dtype=torch.int32, device=coords.device)
    return float(tour_lengths(coords[None], t, n)[0])
