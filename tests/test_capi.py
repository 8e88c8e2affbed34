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
