"""ctypes binding of libhtsp_b200.so (include/htsp_b200.h).  The library is the product: if
it is missing or fails to load this module raises -- there is no CPU or PyTorch fallback."""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# HTSP_LIB_PATH: another build of the same library (tools/ab_rollout.py compares two builds); never a fallback
LIB_PATH = os.environ.get("HTSP_LIB_PATH") or os.path.join(_PKG_DIR, "libhtsp_b200.so")

MODE_FP32, MODE_TC_X3, MODE_TC_FAST = 0, 1, 2
MODES = {"fp32": MODE_FP32, "x3": MODE_TC_X3, "fast": MODE_TC_FAST}

_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t

# name -> (restype, argtypes); mirrors include/htsp_b200.h one to one
SIGNATURES = {
    "htsp_abi_version": (_i, []),
    "htsp_last_error": (C.c_char_p, []),
    "htsp_launch_count": (C.c_uint64, []),
    "htsp_set_kernel_timer": (_i, [_vp, _vp]),
    "htsp_weights_blob_bytes": (_sz, [_i]),
    "htsp_n_weight_tensors": (_i, [_i]),
    "htsp_pack_weights": (_i, [C.POINTER(_vp), _i, _i, _vp, _vp]),
    "htsp_extract_normalize": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "htsp_augment8": (_i, [_vp, _i, _i, _vp, _vp]),
    "htsp_encode_workspace_bytes": (_sz, [_i, _i, _i]),
    "htsp_encode": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "htsp_encode_train_workspace_bytes": (_sz, [_i, _i, _i]),
    "htsp_encode_train": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "htsp_decode_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "htsp_decode": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp,
                         _vp, _sz, _vp]),
    "htsp_decode_sample": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, C.c_uint64, _vp, _vp, _vp, _sz,
                                _vp]),
    "htsp_decode_sample_logp": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, C.c_uint64, _vp, _vp, _vp, _vp,
                                     _sz, _vp]),
    "htsp_decode_ctas_per_sm": (_i, [_i, _i, _i, _i]),
    "htsp_decode_key_blocks": (_i, [_i, _vp]),
    "htsp_decode_key_block_stage": (_i, [_i, _i, _vp]),
    "htsp_decode_debug_floats": (_sz, [_i]),
    "htsp_decode_debug": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp,
                               _vp, _sz, _vp, _i, _vp]),
    "htsp_select_best": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "htsp_finalize_paths": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "htsp_tour_length": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "htsp_env_knn": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "htsp_env_fragment": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp,
                               _vp, _vp, _vp]),
    "htsp_env_splice": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "htsp_policy_n_tensors": (_i, []),
    "htsp_policy_blob_bytes": (_sz, [_i, _i, _i]),
    "htsp_policy_pack_weights": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _vp, _vp]),
    "htsp_policy_workspace_bytes": (_sz, [_i, _i, _i]),
    "htsp_policy_forward": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "htsp_env_update": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp,
                             _vp, _vp]),
}

_lib = None


class HtspError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise HtspError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  htsp_b200 has no fallback path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if handle.htsp_abi_version() != 1:
            raise HtspError("libhtsp_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise HtspError(f"{what} failed ({rc}): {lib().htsp_last_error().decode(errors='replace')}")


def ptr(t) -> int:
    """Device (or host) address of a torch tensor, None -> NULL."""
    return None if t is None else t.data_ptr()
