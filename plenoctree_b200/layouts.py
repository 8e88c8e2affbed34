"""numpy mirrors of the shared-memory operand layouts and tcgen05 descriptors of csrc/common.cuh.

Used by the tests to build operand images for the UMMA probe and to check the device pack
kernel; kept next to the product code because the layouts are part of the kernel contract.
"""
import numpy as np

TILE_M = 128
A_CHUNK_BYTES = 128 * 128
WSLOT_BYTES = 256 * 64

LAYOUT_NONE, LAYOUT_SW128, LAYOUT_SW64, LAYOUT_SW32 = 0, 2, 4, 6


def a_tile_offset(row, col):
    """byte offset of fp16 element (row, col) in a K-major SW128 activation tile image."""
    row = np.asarray(row)
    col = np.asarray(col)
    return ((col >> 6) * A_CHUNK_BYTES + row * 128 + ((((col >> 3) & 7) ^ (row & 7)) << 4)
            + (col & 7) * 2)


def w_slot_offset(row, k):
    """byte offset of fp16 element (row, k<32) in a K-major SW64 weight slot image."""
    row = np.asarray(row)
    k = np.asarray(k)
    return row * 64 + ((((k >> 3) & 3) ^ ((row >> 1) & 3)) << 4) + (k & 7) * 2


def pack_a_tile(mat):
    """[128, 64*n] float -> uint8 image (fp16, SW128)."""
    rows, cols = mat.shape
    assert rows == 128 and cols % 64 == 0
    img = np.zeros(A_CHUNK_BYTES * (cols // 64), dtype=np.uint8)
    h = mat.astype(np.float16).view(np.uint16)
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    off = a_tile_offset(r, c)
    img16 = img.view(np.uint16)
    img16[off // 2] = h
    return img


def unpack_a_tile(img, cols):
    r, c = np.meshgrid(np.arange(128), np.arange(cols), indexing="ij")
    off = a_tile_offset(r, c)
    return img.view(np.uint16)[off // 2].view(np.float16).astype(np.float32)


def t_tile_offset(row, col):
    """byte offset of fp16 element (row, col) in a "T" tile image: [32-row group][8-column unit]
    [row in group][16 B].  mlp_fwd writes the saved h_l tiles in this order (one coalesced 512 B
    store per warp instruction straight from the epilogue registers); read MN-major by the tensor
    cores it is the canonical no-swizzle layout with 128 B core matrices (8 rows x 8 columns):
    LBO (next 8 rows) = 128 B, SBO (next 8 columns) = 512 B."""
    row = np.asarray(row)
    col = np.asarray(col)
    return (row >> 5) * 16384 + (col >> 3) * 512 + (row & 31) * 16 + (col & 7) * 2


def pack_t_tile(mat):
    """[128, 256] float -> uint8 image (fp16, T layout)."""
    rows, cols = mat.shape
    assert rows == 128 and cols == 256
    img = np.zeros(rows * cols * 2, dtype=np.uint8)
    r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
    img.view(np.uint16)[t_tile_offset(r, c) // 2] = mat.astype(np.float16).view(np.uint16)
    return img


def unpack_t_tile(img):
    r, c = np.meshgrid(np.arange(128), np.arange(256), indexing="ij")
    return img.view(np.uint16)[t_tile_offset(r, c) // 2].view(np.float16).astype(np.float32)


def pack_w_slot(mat):
    """[rows, 32] float -> uint8 image (fp16, SW64)."""
    rows, k = mat.shape
    assert k == 32 and rows % 8 == 0
    img = np.zeros(rows * 64, dtype=np.uint8)
    r, c = np.meshgrid(np.arange(rows), np.arange(32), indexing="ij")
    img.view(np.uint16)[w_slot_offset(r, c) // 2] = mat.astype(np.float16).view(np.uint16)
    return img


def make_idesc_f16(M, N, a_mn_major=0, b_mn_major=0):
    return (1 << 4) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def make_sdesc(addr, lbo_bytes, sbo_bytes, layout):
    return (((addr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16)
            | (((sbo_bytes >> 4) & 0x3FFF) << 32) | (1 << 46) | (layout << 61))


# ---- flat parameter layout / packed image geometry (mirrors csrc/kernels.h, pack.cu) ---------
def K_of(sh_deg):
    return 1 if sh_deg < 0 else (sh_deg + 1) ** 2


def heads_width(K):
    return (1 + 3 * K + 15) // 16 * 16


def fwd_has_bias_slot(l):
    return not (l == 0 or l == 5)


def fwd_slots_of_layer(l):
    return 2 if l == 0 else (10 if l == 5 else 9)


def layer_dims(K):
    dims = []
    for i in range(10):
        cin = 63 if i == 0 else (319 if i == 5 else 256)
        cout = 256 if i < 8 else (1 if i == 8 else 3 * K)
        dims.append((cin, cout))
    return dims


def flat_offsets(K):
    w_off, b_off, off = [], [], 0
    for cin, cout in layer_dims(K):
        w_off.append(off)
        off += cin * cout
        b_off.append(off)
        off += cout
    return w_off, b_off, off


def blob_layout(K):
    NH = heads_width(K)
    up = lambda x: (x + 1023) // 1024 * 1024
    fwd = 66 * 16384 + 9 * NH * 64
    bwd = ((NH + 31) // 32 + 56) * 16384
    w_hi = 0
    w_lo = up(w_hi + fwd)
    wt_hi = up(w_lo + fwd)
    bias = up(wt_hi + bwd)
    total = up(bias + (8 * 256 + 80) * 4)
    return dict(w_hi=w_hi, w_lo=w_lo, wt_hi=wt_hi, bias=bias, total=total, fwd_bytes=fwd,
                bwd_bytes=bwd, NH=NH)


def heads_matrix(flat, K):
    """packed heads weight [256 in, NH] and bias [NH] in kernel column order [sigma, (k,c)...]."""
    w_off, b_off, _ = flat_offsets(K)
    NH = heads_width(K)
    W8 = flat[w_off[8]:w_off[8] + 256].reshape(256, 1)
    W9 = flat[w_off[9]:w_off[9] + 256 * 3 * K].reshape(256, 3 * K)
    b8 = flat[b_off[8]:b_off[8] + 1]
    b9 = flat[b_off[9]:b_off[9] + 3 * K]
    Wh = np.zeros((256, NH), np.float32)
    bh = np.zeros(NH, np.float32)
    Wh[:, 0] = W8[:, 0]
    bh[0] = b8[0]
    for k in range(K):
        for c in range(3):
            Wh[:, 1 + 3 * k + c] = W9[:, c * K + k]
            bh[1 + 3 * k + c] = b9[c * K + k]
    return Wh, bh


def pack_reference(flat, sh_deg):
    """numpy model of pack.cu: returns dict of uint8 images w_hi, w_lo, wt_hi and float32 bias."""
    K = K_of(sh_deg)
    L = blob_layout(K)
    NH = L["NH"]
    w_off, b_off, total = flat_offsets(K)
    flat = np.asarray(flat, np.float32)
    assert flat.size == total
    dims = layer_dims(K)
    w_hi = np.zeros(L["fwd_bytes"], np.uint8)
    w_lo = np.zeros(L["fwd_bytes"], np.uint8)
    wt_hi = np.zeros(L["bwd_bytes"], np.uint8)

    def hilo(m):
        hi = m.astype(np.float16)
        lo = (m - hi.astype(np.float32)).astype(np.float16)
        return hi, lo

    slot = 0
    for l in range(8):
        cin = dims[l][0]
        W = flat[w_off[l]:w_off[l] + cin * 256].reshape(cin, 256)  # [in, out]
        bias_l = flat[b_off[l]:b_off[l] + 256]
        for j in range(fwd_slots_of_layer(l)):
            blk = np.zeros((256, 32), np.float32)  # [out row, k]
            if fwd_has_bias_slot(l) and j == 8:
                blk[:, 31] = bias_l                # bias slot: multiplied by posenc column 63 (= 1)
            else:
                k0 = 32 * j
                kn = max(0, min(32, cin - k0))
                blk[:, :kn] = W[k0:k0 + kn, :].T
                if not fwd_has_bias_slot(l) and k0 <= cin < k0 + 32:
                    blk[:, cin - k0] = bias_l      # layers 0 / 5: padding row k = 63 of the posenc operand
            hi, lo = hilo(blk)
            w_hi[slot * 16384:(slot + 1) * 16384] = pack_w_slot(hi)
            w_lo[slot * 16384:(slot + 1) * 16384] = pack_w_slot(lo)
            slot += 1
    Wh, bh = heads_matrix(flat, K)
    base = 66 * 16384
    for j in range(9):
        if j < 8:
            blk = Wh[32 * j:32 * j + 32, :].T  # [NH, 32]
        else:
            blk = np.zeros((NH, 32), np.float32)
            blk[:, 31] = bh
        hi, lo = hilo(blk)
        w_hi[base + j * NH * 64: base + (j + 1) * NH * 64] = pack_w_slot(hi)
        w_lo[base + j * NH * 64: base + (j + 1) * NH * 64] = pack_w_slot(lo)
    # dgrad images: rows = in feature, k = out feature
    hs = (NH + 31) // 32
    slot = 0
    for j in range(hs):
        blk = np.zeros((256, 32), np.float32)
        kn = max(0, min(32, NH - 32 * j))
        blk[:, :kn] = Wh[:, 32 * j:32 * j + kn]
        wt_hi[slot * 16384:(slot + 1) * 16384] = pack_w_slot(blk)
        slot += 1
    for l in range(7, 0, -1):
        cin = dims[l][0]
        W = flat[w_off[l]:w_off[l] + cin * 256].reshape(cin, 256)[:256]  # [in(256), out]
        for j in range(8):
            wt_hi[slot * 16384:(slot + 1) * 16384] = pack_w_slot(W[:, 32 * j:32 * j + 32])
            slot += 1
    bias = np.zeros(8 * 256 + 80, np.float32)
    for l in range(8):
        bias[l * 256:(l + 1) * 256] = flat[b_off[l]:b_off[l] + 256]
    bias[2048:2048 + NH] = bh
    return dict(w_hi=w_hi, w_lo=w_lo, wt_hi=wt_hi, bias=bias)
