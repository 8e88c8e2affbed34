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
