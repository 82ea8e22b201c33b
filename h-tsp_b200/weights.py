"""Weight handling for the path solver: the reference state_dict key layout
(rl4cop/models/models.py:12-60,331-358) and a seeded synthetic
generator used by tests and bench.py (trained checkpoints are Google-Drive links that are
not available offline, reference README.md:94)."""
from __future__ import annotations

import math
from typing import Dict

import torch

EMBED_DIM = 128
N_HEADS = 8
FF_DIM = 512


def expected_keys(n_layers: int):
    keys = ["encoder.init_projection_layer.weight", "encoder.init_projection_layer.bias"]
    for i in range(n_layers):
        p = f"encoder.attn_layers.{i}."
        keys += [p + "Wq.weight", p + "Wk.weight", p + "Wv.weight",
                 p + "multi_head_combine.weight", p + "multi_head_combine.bias",
                 p + "feed_forward.0.weight", p + "feed_forward.0.bias",
                 p + "feed_forward.2.weight", p + "feed_forward.2.bias"]
        for n in ("norm1", "norm2"):
            keys += [p + f"{n}.weight", p + f"{n}.bias", p + f"{n}.running_mean",
                     p + f"{n}.running_var"]
    keys += [f"decoder.{w}.weight" for w in
             ("Wq_graph", "Wq_source", "Wq_target", "Wq_first", "Wq_last", "Wk", "Wv")]
    keys += ["decoder.multi_head_combine.weight", "decoder.multi_head_combine.bias"]
    return keys


def n_layers_of(state_dict) -> int:
    n = 0
    while f"encoder.attn_layers.{n}.Wq.weight" in state_dict:
        n += 1
    return n


def synthetic_state_dict(seed: int = 1234, n_layers: int = 6) -> Dict[str, torch.Tensor]:
    """Random-init weights with the distribution torch.nn.Linear uses by default
    (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias) and NON-trivial BatchNorm
    statistics (weight~U(0.5,1.5), bias~N(0,0.1), running_mean~N(0,0.1),
    running_var~U(0.5,1.5)) so that eval-mode BN is exercised.  Deterministic in (seed,
    n_layers) for a given torch build; fp32, CPU."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    H, F = EMBED_DIM, FF_DIM

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    sd: Dict[str, torch.Tensor] = {}
    sd["encoder.init_projection_layer.weight"] = uni((H, 2), 2)
    sd["encoder.init_projection_layer.bias"] = uni((H,), 2)
    for i in range(n_layers):
        p = f"encoder.attn_layers.{i}."
        for w in ("Wq", "Wk", "Wv"):
            sd[p + w + ".weight"] = uni((H, H), H)
        sd[p + "multi_head_combine.weight"] = uni((H, H), H)
        sd[p + "multi_head_combine.bias"] = uni((H,), H)
        sd[p + "feed_forward.0.weight"] = uni((F, H), H)
        sd[p + "feed_forward.0.bias"] = uni((F,), H)
        sd[p + "feed_forward.2.weight"] = uni((H, F), F)
        sd[p + "feed_forward.2.bias"] = uni((H,), F)
        for n in ("norm1", "norm2"):
            sd[p + f"{n}.weight"] = torch.rand(H, generator=g) + 0.5
            sd[p + f"{n}.bias"] = torch.randn(H, generator=g) * 0.1
            sd[p + f"{n}.running_mean"] = torch.randn(H, generator=g) * 0.1
            sd[p + f"{n}.running_var"] = torch.rand(H, generator=g) + 0.5
    for w in ("Wq_graph", "Wq_source", "Wq_target", "Wq_first", "Wq_last", "Wk", "Wv"):
        sd[f"decoder.{w}.weight"] = uni((H, H), H)
    sd["decoder.multi_head_combine.weight"] = uni((H, H), H)
    sd["decoder.multi_head_combine.bias"] = uni((H,), H)
    return sd
