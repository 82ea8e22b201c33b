"""TEST INFRASTRUCTURE ONLY -- records the upper-level policy forward of the UNMODIFIED reference
(rl_models.IMPALAEncoder + rl_models.ActorPPO, as HTSP_PPO.forward composes them, h_tsp.py:1272-1276)
into tests/golden/upper_policy_golden.npz.  Run in the build container:
    python oracle/make_golden_policy.py

pytorch-scatter is absent here; oracle/ref_loader.py restates its two functions (see there).  Weights are
`synthetic_policy_state_dict(seed)` (numpy RandomState: the same bits everywhere) loaded with
strict=True into the reference modules, so only inputs and outputs are stored.  Inputs are environment
states recorded from reference episodes (tests/golden/env_state_golden.npz) plus a clustered instance
with many points per pillar."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

SEED = 4321


def clustered_state(rs, B, N):
    """A state tensor with heavy pillar sharing: Gaussian blobs, clipped into [0, 1)."""
    s = np.zeros((B, N, 8), dtype=np.float32)
    for b in range(B):
        centres = rs.uniform(0.15, 0.85, size=(5, 2))
        xy = centres[rs.randint(0, 5, size=N)] + rs.normal(0, 0.03, size=(N, 2))
        s[b, :, 0:2] = np.clip(xy, 0.0, 0.999).astype(np.float32)
        s[b, :, 2] = rs.rand(N) < 0.5
        s[b, :, 3] = (rs.rand(N) < 0.3) * s[b, :, 2]
        nb = s[b, rs.randint(0, N, size=(N, 2)), 0:2].reshape(N, 4)
        s[b, :, 4:8] = nb * s[b, :, 2:3]
    return s


def main():
    M = ref_loader.load_rl_models_module()
    pol = importlib.import_module("h-tsp_b200.upper_policy")
    sd = pol.synthetic_policy_state_dict(SEED)
    enc = M.IMPALAEncoder(input_dim=8, embedding_dim=128)           # h_tsp.py:949-952
    act = M.ActorPPO(state_dim=128, mid_dim=128, action_dim=2, init_a_std_log=-1)   # h_tsp.py:959-964
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    missing, unexpected = act.load_state_dict({k[len("actor."):]: v for k, v in sd.items() if k.startswith("actor.")},
                                              strict=False)
    assert missing == ["a_std_log"] and not unexpected, (missing, unexpected)
    enc.eval()
    act.eval()

    env = np.load(os.path.join(ROOT, "tests", "golden", "env_state_golden.npz"))
    rs = np.random.RandomState(99)
    cases = {
        "n300": env["n300_states"][[0, 5, 16]],             # early / mid / finished construction
        "n1000": env["n1000_states"],
        "cluster": clustered_state(rs, 2, 2000),
    }
    # checksum of the generated weights: detects any drift of the generator between boxes
    out = {"seed": np.int64(SEED),
           "weights_checksum": np.float64(sum(float(v.double().abs().sum()) for v in sd.values()))}
    for tag, s in cases.items():
        st = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32))
        taps = {}
        h = enc.cnn.register_forward_pre_hook(lambda m, inp: taps.__setitem__("img", inp[0].detach().clone()))
        with torch.no_grad():
            feat = enc(st)
            a = act(feat)
        h.remove()
        img = taps["img"].numpy()
        out[f"{tag}_state"] = st.numpy()
        out[f"{tag}_img_nonzero"] = np.int64((img != 0).sum())
        out[f"{tag}_img_chan_sum"] = img.astype(np.float64).sum(axis=(2, 3))           # [B,C0]
        out[f"{tag}_img_rows"] = img[:, :, ::7, :].max(axis=3)                         # [B,C0,15] row maxima
        out[f"{tag}_feat"] = feat.numpy()
        out[f"{tag}_action"] = a.numpy()
        print(tag, st.shape, "nonzero pixels", int((img != 0).sum()), "action", a.numpy().round(4).tolist())
    path = os.path.join(ROOT, "tests", "golden", "upper_policy_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
