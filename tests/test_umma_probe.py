"""Pins the tcgen05 descriptor conventions the production kernels use, on a real B200, with the
single-CTA probe kernel (csrc/probe.cu): operand images are built with the numpy layout mirrors and
the accumulator is compared with a numpy matmul of the fp16-rounded operands (fp32 accumulate)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _record(name, payload):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "umma_probe.json")
    data = {}
    if os.path.exists(path):
        try:
            data = json.load(open(path))
        except Exception:
            data = {}
    data[name] = payload
    json.dump(data, open(path, "w"), indent=1)


def _f16(a):
    return a.astype(np.float16).astype(np.float32)


def _relerr(got, want):
    return float(np.abs(got - want).max() / max(1e-6, np.abs(want).max()))


def test_kmajor_sw128_a_times_sw64_weight_slots():
    """forward / dgrad operand pair: A = activation tile (K-major SW128), B = weight slots
    (K-major SW64, 32 k each), k16 sub-steps by +32 B on the start address."""
    from plenoctree_b200 import layouts as L, ops
    rs = np.random.RandomState(0)
    for N in (256, 64, 80):
        A = _f16(rs.normal(size=(128, 64)))
        B = _f16(rs.normal(size=(N, 64)))
        a_img = L.pack_a_tile(A)
        b_img = np.concatenate([L.pack_w_slot(B[:, 0:32]), L.pack_w_slot(B[:, 32:64])])
        b_off = 16384
        slot_bytes = N * 64
        ad, bd, dc, ac = [], [], [], []
        for j in range(2):
            for ks in range(2):
                ad.append(L.make_sdesc(j * 64 + ks * 32, 16, 1024, L.LAYOUT_SW128))
                bd.append(L.make_sdesc(b_off + j * slot_bytes + ks * 32, 16, 512, L.LAYOUT_SW64))
                dc.append(0)
                ac.append(0 if (j | ks) == 0 else 1)
        got = ops.umma_probe(a_img, b_img, b_off, ad, bd, dc, ac, L.make_idesc_f16(128, N), N)
        err = _relerr(got, A @ B.T)
        _record(f"kmajor_N{N}", err)
        assert err < 1e-5, f"N={N} rel err {err}"


def test_kmajor_multi_chunk_k256():
    """K = 256 across four SW128 chunks of the activation tile and eight weight slots."""
    from plenoctree_b200 import layouts as L, ops
    rs = np.random.RandomState(1)
    A = _f16(rs.normal(size=(128, 256)))
    B = _f16(rs.normal(size=(256, 256)))
    a_img = L.pack_a_tile(A)
    b_img = np.concatenate([L.pack_w_slot(B[:, 32 * j:32 * j + 32]) for j in range(8)])
    b_off = 65536
    ad, bd, dc, ac = [], [], [], []
    for j in range(8):
        for ks in range(2):
            ad.append(L.make_sdesc((j >> 1) * L.A_CHUNK_BYTES + (j & 1) * 64 + ks * 32, 16, 1024, L.LAYOUT_SW128))
            bd.append(L.make_sdesc(b_off + j * 16384 + ks * 32, 16, 512, L.LAYOUT_SW64))
            dc.append(256)  # second accumulator half, like tile Y
            ac.append(0 if (j | ks) == 0 else 1)
    got = ops.umma_probe(a_img, b_img, b_off, ad, bd, dc, ac, L.make_idesc_f16(128, 256), 512)
    err = _relerr(got[:, 256:], A @ B.T)
    _record("kmajor_k256", err)
    assert err < 1e-5


def _mn_major_case(lbo, sbo):
    """weight-gradient contraction: D[f_out, f_in] = dZ[:, f_out]^T . H  with both operands read
    MN-major from [sample x feature] activation tile images."""
    from plenoctree_b200 import layouts as L, ops
    rs = np.random.RandomState(2)
    DZ = _f16(rs.normal(size=(128, 256)))
    H = _f16(rs.normal(size=(128, 256)))
    a_img = L.pack_a_tile(DZ)
    b_img = L.pack_a_tile(H)
    b_off = 65536
    ad, bd, dc, ac = [], [], [], []
    for half in range(2):
        for ks in range(8):  # 16 samples per MMA
            ad.append(L.make_sdesc(half * 2 * L.A_CHUNK_BYTES + ks * 2048, lbo, sbo, L.LAYOUT_SW128))
            bd.append(L.make_sdesc(b_off + ks * 2048, lbo, sbo, L.LAYOUT_SW128))
            dc.append(256 * half)
            ac.append(0 if ks == 0 else 1)
    idesc = L.make_idesc_f16(128, 256, a_mn_major=1, b_mn_major=1)
    got = ops.umma_probe(a_img, b_img, b_off, ad, bd, dc, ac, idesc, 512)
    want = DZ.T @ H  # [256 f_out, 256 f_in]
    e0 = _relerr(got[:, :256], want[:128])
    e1 = _relerr(got[:, 256:], want[128:])
    return max(e0, e1)


def test_mn_major_sw128_wgrad_operands():
    # convention used by mlp_wgrad: LBO = stride between 64-feature chunks, SBO = 8-sample group
    errs = {}
    for name, (lbo, sbo) in {"lbo_chunk_sbo_1024": (16384, 1024), "swapped": (1024, 16384)}.items():
        errs[name] = _mn_major_case(lbo, sbo)
    _record("mn_major", errs)
    assert errs["lbo_chunk_sbo_1024"] < 1e-5, errs


def _t_layout_case(lbo, sbo, h_is_b):
    """weight-gradient contraction with the forward-saved h tile in the T layout (no swizzle, MN-major)
    on one side and a SW128 dZ tile on the other, 64 samples per stage like mlp_wgrad."""
    from plenoctree_b200 import layouts as L, ops
    rs = np.random.RandomState(3)
    DZ = _f16(rs.normal(size=(128, 256)))
    H = _f16(rs.normal(size=(128, 256)))
    sw_img = L.pack_a_tile(DZ)
    t_img = L.pack_t_tile(H)
    off2 = 65536
    ad, bd, dc, ac = [], [], [], []
    for half in range(2):
        for ks in range(8):  # 16 samples per MMA
            sw = L.make_sdesc(half * 2 * L.A_CHUNK_BYTES + ks * 2048, 16384, 1024, L.LAYOUT_SW128)
            # T image: 32-sample group ks>>1 (16 KB), 16-sample half (ks&1) = two 128 B core matrices
            tt = L.make_sdesc(off2 + (ks >> 1) * 16384 + (ks & 1) * 256, lbo, sbo, L.LAYOUT_NONE)
            if h_is_b:
                ad.append(sw)
                bd.append(tt)
            else:   # heads role: A = h (features 128*half ..), B = the SW128 tile
                ad.append(L.make_sdesc(off2 + (ks >> 1) * 16384 + (ks & 1) * 256 + half * 16 * sbo, lbo, sbo, L.LAYOUT_NONE) if sbo == 512
                          else L.make_sdesc(off2 + (ks >> 1) * 16384 + (ks & 1) * 256 + half * 16 * lbo, lbo, sbo, L.LAYOUT_NONE))
                bd.append(L.make_sdesc(ks * 2048, 16384, 1024, L.LAYOUT_SW128))
            dc.append(256 * half)
            ac.append(0 if ks == 0 else 1)
    idesc = L.make_idesc_f16(128, 256, a_mn_major=1, b_mn_major=1)
    got = ops.umma_probe(sw_img, t_img, off2, ad, bd, dc, ac, idesc, 512)
    want = (DZ.T @ H) if h_is_b else (H.T @ DZ)
    return max(_relerr(got[:, :256], want[:128]), _relerr(got[:, 256:], want[128:]))


def test_mn_major_t_layout_wgrad_operands():
    errs = {}
    for name, (lbo, sbo) in {"lbo128_sbo512": (128, 512), "lbo512_sbo128": (512, 128)}.items():
        errs[name + "_b"] = _t_layout_case(lbo, sbo, True)
        errs[name + "_a"] = _t_layout_case(lbo, sbo, False)
    _record("mn_major_t_layout", errs)
    ok = [k[:-2] for k in errs if errs[k] < 1e-5]
    assert any(ok.count(n) == 2 for n in ok), errs


def test_cta_pair_kmajor_weight_halves():
    """cta_group::2 (M = 256 over a CTA pair): each CTA stages its own 128 activation rows (K-major SW128)
    and HALF of the weight rows of every slot (K-major SW64: rows 128r..128r+127 = bytes [8192r, +8192) of
    a 16 KB slot); both CTAs receive all N accumulator columns for their rows."""
    from plenoctree_b200 import layouts as L, ops
    rs = np.random.RandomState(4)
    for N in (256, 64, 80):
        A = _f16(rs.normal(size=(256, 64)))
        B = _f16(rs.normal(size=(N, 64)))
        a_img = np.stack([L.pack_a_tile(A[:128]), L.pack_a_tile(A[128:])])
        hb = N // 2 * 64  # bytes of one CTA's half of a slot
        slots = [L.pack_w_slot(B[:, 0:32]), L.pack_w_slot(B[:, 32:64])]
        b_img = np.stack([np.concatenate([s[r * hb:(r + 1) * hb] for s in slots]) for r in range(2)])
        b_off = 16384
        ad, bd, dc, ac = [], [], [], []
        for j in range(2):
            for ks in range(2):
                ad.append(L.make_sdesc(j * 64 + ks * 32, 16, 1024, L.LAYOUT_SW128))
                bd.append(L.make_sdesc(b_off + j * hb + ks * 32, 16, 512, L.LAYOUT_SW64))
                dc.append(0)
                ac.append(0 if (j | ks) == 0 else 1)
        got = ops.umma_probe(a_img, b_img, b_off, ad, bd, dc, ac, L.make_idesc_f16(256, N), N, pair=True)
        err = _relerr(got, A @ B.T)
        _record(f"pair_kmajor_N{N}", err)
        assert err < 1e-5, f"N={N} rel err {err}"
