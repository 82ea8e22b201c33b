"""B200UpperPolicy: the upper-level policy forward of the reference,
`HTSP_PPO.forward` (h_tsp.py:1272-1276) = `ActorPPO.forward` (rl_models.py:329-336) o
`IMPALAEncoder.forward` (rl_models.py:187-273), as kernels of libhtsp_b200 (csrc/upper_policy.cu)
behind `htsp_policy_forward`: environment state [E,Ntot,8] in HBM (the tensor `htsp_env_update`
writes) -> actions [E,2] in HBM, so that a whole `VecEnv` step stays on the device.

The reference needs pytorch-scatter and `torch.unique` for the pillar stage; here pillars are addressed
by their pixel id directly.  torch provides device memory and streams only; there is no CPU path."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _abi

INPUT_DIM = 8            # config_ppo.yaml:15; state = [x, y, selected, available, neighbour coords(4)]
GRID = 100               # bev_range [0,0,1,1] / bev_pixel_size 0.01 (rl_models.py:155-158)
DEPTHS = (16, 32, 32)    # IMPALACNN default depths (rl_models.py:29-30)


def _flat_hw() -> int:
    s = GRID
    for _ in DEPTHS:
        s = (s + 2 - 3) // 2 + 1          # MaxPool2d(kernel 3, stride 2, padding 1)
    return s


FLAT_DIM = DEPTHS[-1] * _flat_hw() ** 2   # 32 * 13 * 13 = 5408


def expected_policy_shapes(embedding_dim: int = 128, mid_dim: int = 128, action_dim: int = 2
                           ) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) of the tensors `htsp_policy_pack_weights` takes, in its order; keys are those of
    `HTSP_PPO.state_dict()` (h_tsp.py:949-964).  `actor.a_std_log` is not used by forward()."""
    c0 = embedding_dim // 2
    out = [("encoder.input_transform.0.weight", (c0, INPUT_DIM + 4, 1)),
           ("encoder.input_transform.1.weight", (c0,)), ("encoder.input_transform.1.bias", (c0,))]
    cin = c0
    for i, d in enumerate(DEPTHS):
        out += [(f"encoder.cnn.feat_convs.{i}.0.weight", (d, cin, 3, 3)), (f"encoder.cnn.feat_convs.{i}.0.bias", (d,))]
        for blk in ("resnet1", "resnet2"):
            for j in (1, 3):
                out += [(f"encoder.cnn.{blk}.{i}.{j}.weight", (d, d, 3, 3)), (f"encoder.cnn.{blk}.{i}.{j}.bias", (d,))]
        cin = d
    out += [("encoder.cnn.fc.weight", (embedding_dim, FLAT_DIM)), ("encoder.cnn.fc.bias", (embedding_dim,))]
    out += [("actor.net.0.0.weight", (mid_dim, embedding_dim)), ("actor.net.0.0.bias", (mid_dim,)),
            ("actor.net.0.2.weight", (mid_dim, mid_dim)), ("actor.net.0.2.bias", (mid_dim,)),
            ("actor.net.1.weight", (mid_dim, mid_dim)), ("actor.net.1.bias", (mid_dim,)),
            ("actor.net.3.weight", (action_dim, mid_dim)), ("actor.net.3.bias", (action_dim,))]
    return out


def synthetic_policy_state_dict(seed: int = 1234, embedding_dim: int = 128, mid_dim: int = 128,
                                action_dim: int = 2) -> Dict[str, torch.Tensor]:
    """Random-init weights (trained checkpoints are not available offline, reference README.md:94) from
    numpy's legacy RandomState, so the same (seed) gives the same bits on every box: weights
    U(-b, b) with b = sqrt(3 / fan_in) (unit-gain), biases U(-0.1, 0.1), InstanceNorm weight U(0.5, 1.5)
    and bias U(-0.3, 0.3); the last actor layer is scaled so that actions spread over (-1, 1)."""
    rs = np.random.RandomState(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in expected_policy_shapes(embedding_dim, mid_dim, action_dim):
        if key == "encoder.input_transform.1.weight":
            v = rs.uniform(0.5, 1.5, size=shape)
        elif key == "encoder.input_transform.1.bias":
            v = rs.uniform(-0.3, 0.3, size=shape)
        elif key.endswith(".bias"):
            v = rs.uniform(-0.1, 0.1, size=shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            gain = 2.0 if "cnn" in key or "actor" in key else 1.0       # ReLU halves the variance
            b = np.sqrt(3.0 * gain / fan_in)
            v = rs.uniform(-b, b, size=shape)
            if key == "actor.net.3.weight":
                v = v * 0.04
        sd[key] = torch.from_numpy(v.astype(np.float32))
    return sd


class B200UpperPolicy:
    """`model(s)` of evaluate.py:84: states [E,Ntot,8] f32 (device) -> actions [E,2] in (-1,1)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device, embedding_dim: int = 128,
                 mid_dim: int = 128, action_dim: int = 2):
        self.device = torch.device(device)
        assert self.device.type == "cuda", "htsp_b200 has no CPU path"
        self.embedding_dim, self.mid_dim, self.action_dim = embedding_dim, mid_dim, action_dim
        lib = _abi.lib()
        shapes = expected_policy_shapes(embedding_dim, mid_dim, action_dim)
        assert lib.htsp_policy_n_tensors() == len(shapes)
        host = []
        for key, shape in shapes:
            if key not in state_dict:
                raise KeyError(f"upper policy: missing weight {key}")
            t = state_dict[key].detach().to("cpu", torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"upper policy: {key} has shape {tuple(t.shape)}, expected {shape}")
            host.append(t)
        nbytes = lib.htsp_policy_blob_bytes(embedding_dim, mid_dim, action_dim)
        if nbytes == 0:
            raise _abi.HtspError("upper policy: unsupported dimensions")
        self._blob = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        arr = (C.c_void_p * len(host))(*[t.data_ptr() for t in host])
        with torch.cuda.device(self.device):
            _abi.check(lib.htsp_policy_pack_weights(arr, len(host), embedding_dim, mid_dim, action_dim,
                                                    self._blob.data_ptr(), self._stream()),
                       "htsp_policy_pack_weights")
        self._ws = None
        self._ws_key = None

    @classmethod
    def from_reference(cls, module, device):
        """module: the reference's HTSP_PPO (or anything whose state_dict has encoder.* / actor.* keys)."""
        sd = module.state_dict()
        e = sd["encoder.cnn.fc.weight"].shape[0]
        return cls(sd, device, embedding_dim=e, mid_dim=sd["actor.net.0.0.weight"].shape[0],
                   action_dim=sd["actor.net.3.weight"].shape[0])

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _workspace(self, E: int, N: int) -> torch.Tensor:
        if self._ws_key != (E, N):
            n = _abi.lib().htsp_policy_workspace_bytes(E, N, self.embedding_dim)
            self._ws = torch.empty(n, dtype=torch.uint8, device=self.device)
            self._ws_key = (E, N)
        return self._ws

    @torch.no_grad()
    def forward(self, states: torch.Tensor, return_features: bool = False, return_image: bool = False):
        assert states.is_cuda and states.ndim == 3 and states.shape[2] == INPUT_DIM, \
            "states must be a [E,Ntot,8] CUDA tensor"                  # rl_models.py:204
        s = states.to(torch.float32).contiguous()
        E, N, _ = s.shape
        ws = self._workspace(E, N)
        feat = torch.empty(E, self.embedding_dim, dtype=torch.float32, device=self.device)
        act = torch.empty(E, self.action_dim, dtype=torch.float32, device=self.device)
        img = (torch.empty(E, self.embedding_dim // 2, GRID, GRID, dtype=torch.float32, device=self.device)
               if return_image else None)
        with torch.cuda.device(self.device):
            _abi.check(_abi.lib().htsp_policy_forward(self._blob.data_ptr(), self.embedding_dim, self.mid_dim,
                                                      self.action_dim, s.data_ptr(), E, N, feat.data_ptr(),
                                                      act.data_ptr(), _abi.ptr(img), ws.data_ptr(), ws.numel(),
                                                      self._stream()), "htsp_policy_forward")
        out = (act,)
        if return_features:
            out += (feat,)
        if return_image:
            out += (img,)
        return out[0] if len(out) == 1 else out

    __call__ = forward
