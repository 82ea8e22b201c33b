"""CPU tests: the C-ABI library builds for sm_100a, loads without a GPU and exports every symbol
that include/htsp_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "htsp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(htsp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = _declared_symbols()
    for name in ("htsp_pack_weights", "htsp_extract_normalize", "htsp_encode", "htsp_decode",
                 "htsp_select_best", "htsp_finalize_paths", "htsp_tour_length"):
        assert name in syms


def test_library_builds_loads_and_exports_all_symbols(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/htsp_b200.h but not exported"
    lib.htsp_abi_version.restype = ctypes.c_int
    assert lib.htsp_abi_version() == 1


def test_python_binding_covers_the_header(built_lib):
    from htsp_b200 import _abi
    assert sorted(_abi.SIGNATURES) == _declared_symbols()
    lib = _abi.lib()
    assert lib.htsp_n_weight_tensors(6) == 2 + 17 * 6 + 9
    # packed blob = fp32 section (all reference parameters, BN folded 4->2 vectors, + the transposed
    # combine) followed by fp16 hi/lo copies of every GEMM weight for the tensor-core layers
    fp32_bytes = (1318912 + 128 * 128) * 4
    gemm_weights = 6 * (3 * 128 * 128 + 128 * 128 + 2 * 512 * 128) + 640 * 128
    assert fp32_bytes % 256 == 0
    assert lib.htsp_weights_blob_bytes(6) == fp32_bytes + 2 * 2 * gemm_weights


def test_library_contains_sm100a_code_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True)
    if out.returncode != 0:
        return  # cuobjdump unavailable: nothing to check
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_argument_validation_without_gpu(built_lib):
    from htsp_b200 import _abi
    lib = _abi.lib()
    # invalid arguments are rejected before any CUDA call is made
    assert lib.htsp_extract_normalize(None, None, 1, 10, 5, 8, None, None, None, None) == -1
    assert b"null pointer" in lib.htsp_last_error()
    assert lib.htsp_encode_workspace_bytes(0, 200, 0) == 0
    assert lib.htsp_decode_workspace_bytes(4, 200, 200, 0) > 4 * 200 * 640 * 4


def test_rollout_operand_stream_geometry(built_lib, monkeypatch):
    """Host logic of the HTSP_MODE_TC_FAST rollout kernel (csrc/decode_tc2.cu), no GPU: for every supported subproblem
    size the keys are cut into 1 - 3 equal blocks that cover them without an empty block, and the ring stages of one
    decode step tile the operand image of an instance exactly once, each within a ring slot."""
    import ctypes as C
    from htsp_b200 import _abi
    monkeypatch.delenv("HTSP_T2_NKB", raising=False)
    lib = _abi.lib()
    g, st = (C.c_int32 * 5)(), (C.c_uint32 * 2)()
    for bad in (2, 32, 577, 1000):
        assert lib.htsp_decode_key_blocks(bad, g) != 0
    for N in list(range(33, 577, 7)) + [48, 49, 192, 200, 208, 209, 384, 385, 500, 576]:
        assert lib.htsp_decode_key_blocks(N, g) == 0, N
        npb, nkb, stages, img, slot = list(g)
        assert npb % 16 == 0 and npb >= 48
        assert (nkb == 1 and npb <= 208 and N <= 208) or (nkb in (2, 3) and npb <= 192 and N > 208)
        assert nkb * npb >= N > (nkb - 1) * npb                       # every block holds real keys
        assert stages == 3 + 6 * (8 * nkb - 1) + 3 + 12 * nkb
        spans = []
        for i in range(stages):
            assert lib.htsp_decode_key_block_stage(N, i, st) == 0
            assert 0 < st[1] <= slot and st[0] % 1024 == 0 and st[1] % 1024 == 0    # swizzle atoms stay aligned
            spans.append((st[0], st[1]))
        assert lib.htsp_decode_key_block_stage(N, stages, st) != 0
        spans.sort()
        pos = 0
        for off, nbytes in spans:
            assert off == pos, (N, off, pos)
            pos += nbytes
        assert pos == img
