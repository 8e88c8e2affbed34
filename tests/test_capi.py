"""The C-ABI library loads and exports every symbol include/plenoctree_b200.h declares.
No compute calls here (they need a GPU)."""
import os
import re

import numpy as np

from plenoctree_b200 import _lib, layouts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "plenoctree_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pob_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    syms = _header_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in the header but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    for s in _lib.SIGNATURES:
        assert s in syms, f"{s} bound in _lib.py but not declared in the header"


def _header_prototypes():
    """{name: [parameter declarations]} of every `int|int64_t|const char* pob_*(...)` prototype in the header."""
    src = open(os.path.join(ROOT, "include", "plenoctree_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    for m in re.finditer(r"\b(pob_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        protos[m.group(1)] = [] if params in ([""], ["void"]) else params
    return protos


def test_ctypes_signatures_match_the_header():
    """Every binding has as many arguments as the prototype, pointers bound as pointers, floats as c_float,
    64-bit integers as 64-bit — a drifted signature would pass garbage through the ABI without any error."""
    import ctypes as C
    protos = _header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    for name, params in protos.items():
        argtypes = _lib.SIGNATURES[name][1]
        assert len(argtypes) == len(params), (name, len(argtypes), params)
        for decl, ct in zip(params, argtypes):
            is_ptr_ct = ct in (C.c_void_p, C.c_char_p) or hasattr(ct, "contents") or (isinstance(ct, type) and issubclass(ct, C._Pointer))
            if "*" in decl or "[" in decl:                       # `const float offset[3]` decays to a pointer
                assert is_ptr_ct, (name, decl, ct)
            elif re.match(r"(const\s+)?float\b", decl):
                assert ct is C.c_float, (name, decl, ct)
            elif re.match(r"(const\s+)?double\b", decl):
                assert ct is C.c_double, (name, decl, ct)
            elif re.match(r"(const\s+)?(u?int64_t|long long|unsigned long long|size_t)\b", decl):
                assert C.sizeof(ct) == 8 and not is_ptr_ct, (name, decl, ct)
            elif re.match(r"(const\s+)?(int|unsigned|int32_t|uint32_t)\b", decl):
                assert C.sizeof(ct) == 4, (name, decl, ct)
            else:
                raise AssertionError(f"{name}: unclassified parameter {decl!r}")


def test_param_and_blob_sizes():
    assert _lib.lib.pob_abi_version() >= 1
    assert _lib.lib.pob_param_count(3) == 505649
    assert _lib.lib.pob_param_count(4) == 512588
    assert _lib.lib.pob_param_count(7) == -1
    for deg in (-1, 0, 1, 2, 3, 4):
        K = layouts.K_of(deg)
        assert _lib.lib.pob_packed_bytes(deg) == layouts.blob_layout(K)["total"]
        assert layouts.flat_offsets(K)[2] == _lib.lib.pob_param_count(deg)


def test_layout_roundtrip():
    rs = np.random.RandomState(0)
    m = rs.normal(size=(128, 256)).astype(np.float16).astype(np.float32)
    img = layouts.pack_a_tile(m)
    assert img.size == 65536
    np.testing.assert_array_equal(layouts.unpack_a_tile(img, 256), m)
    # every byte of the image written exactly once (offsets are a bijection)
    r, c = np.meshgrid(np.arange(128), np.arange(256), indexing="ij")
    assert np.unique(layouts.a_tile_offset(r, c)).size == 128 * 256
    r, c = np.meshgrid(np.arange(256), np.arange(32), indexing="ij")
    assert np.unique(layouts.w_slot_offset(r, c)).size == 256 * 32


def test_pack_reference_shapes():
    from oracle import nerf_sh_oracle as O
    flat = O.init_flat_params(3, 5, bias_scale=0.1)
    pk = layouts.pack_reference(flat, 3)
    L = layouts.blob_layout(16)
    assert pk["w_hi"].size == L["fwd_bytes"] and pk["wt_hi"].size == L["bwd_bytes"]
    # hi + lo reproduces fp32 weights to ~2^-22 relative
    hi = pk["w_hi"].view(np.float16).astype(np.float64)
    lo = pk["w_lo"].view(np.float16).astype(np.float64)
    assert np.abs(hi).max() > 0.05 and np.abs(lo).max() < 1e-3


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    assert _lib.lib.pob_sm_count() <= 0
    rc = _lib.lib.pob_eval_points_raw(1, 3, 1, 16, None, 1, 1, None)
    assert rc != 0 and len(_lib.lib.pob_last_error()) > 0
