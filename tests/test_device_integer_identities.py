"""CPU models of the integer identities the CUDA kernels rely on (no GPU needed).

The kernels pack, divide and count with tricks whose correctness is pure integer arithmetic; each is restated here in numpy
and checked against the plain definition over the whole range the kernel can see.  (The kernels themselves are checked bit
for bit against the oracle by the -m gpu tests; these tests pin the *reason* they are right.)
"""
import numpy as np


def _vminu2(a, b):
    return (np.minimum(a >> 16, b >> 16) << 16) | np.minimum(a & 0xFFFF, b & 0xFFFF)


def _vmaxu2(a, b):
    return (np.maximum(a >> 16, b >> 16) << 16) | np.maximum(a & 0xFFFF, b & 0xFFFF)


def _fast_score_definition(v, r):
    """cv::FAST-9/16 corner score S' = max over the 16 arcs of nine of min(v - ring) / min(ring - v) (orb.cu: fast_full)."""
    d = v[:, None] - r
    best = np.full(len(v), -10 ** 9)
    for k in range(16):
        idx = [(k + i) & 15 for i in range(9)]
        best = np.maximum(best, np.maximum(d[:, idx].min(1), (-d[:, idx]).min(1)))
    return best


def test_fast_offset_packed_network_equals_the_definition():
    rng = np.random.default_rng(1)
    n = 200000
    v = rng.integers(0, 256, n).astype(np.int64)
    r = np.clip(v[:, None] + rng.integers(-60, 61, (n, 16)), 0, 255).astype(np.int64)
    r[: n // 4] = rng.integers(0, 256, (n // 4, 16))
    r[n // 4: n // 4 + 1000] = 0          # extremes: the halves must never borrow / overflow
    v[n // 4: n // 4 + 500] = 255
    r[n // 4 + 1000: n // 4 + 2000] = 255
    v[n // 4 + 1000: n // 4 + 1500] = 0
    c = ((v + 256) + ((256 - v) << 16)) & 0xFFFFFFFF
    q = [(r[:, k] * 65535 + c) & 0xFFFFFFFF for k in range(16)]
    for k in range(16):  # low half = (v - r) + 256, high half = (r - v) + 256
        assert np.array_equal(q[k] & 0xFFFF, v - r[:, k] + 256) and np.array_equal(q[k] >> 16, r[:, k] - v + 256)
    q3 = [_vminu2(_vminu2(q[k], q[(k + 1) & 15]), q[(k + 2) & 15]) for k in range(16)]
    a9 = [_vminu2(_vminu2(q3[k], q3[(k + 3) & 15]), q3[(k + 6) & 15]) for k in range(16)]
    b = a9[0]
    for k in range(1, 16):
        b = _vmaxu2(b, a9[k])
    assert np.array_equal(np.maximum(b & 0xFFFF, b >> 16) - 256, _fast_score_definition(v, r))


def test_fast_packed_four_point_test_equals_the_definition():
    rng = np.random.default_rng(2)
    n = 200000
    v = rng.integers(0, 256, n).astype(np.int64)
    r = np.clip(v[:, None] + rng.integers(-40, 41, (n, 4)), 0, 255).astype(np.int64)
    r[: n // 8] = rng.integers(0, 256, (n // 8, 4))
    for th in (0, 1, 7, 20, 100, 254, 255):
        k = 0x2000 - 1 - th
        c = ((v + k) + ((k - v) << 16)) & 0xFFFFFFFF
        s = sum((((r[:, j] * 65535 + c) & 0xFFFFFFFF) & 0x20002000) for j in range(4))
        mine = (s & 0xC000C000) != 0
        nb = sum(((v - r[:, j]) > th).astype(int) for j in range(4))
        nd = sum(((r[:, j] - v) > th).astype(int) for j in range(4))
        assert np.array_equal(mine, (nb >= 2) | (nd >= 2)), th


def test_nibble_expansion_and_hamming_through_inner_products():
    """getrt.cu: hm_expand spreads a nibble into four 0/1 bytes; Hamming(a, b) = popc(a) + popc(b) - 2 <a, b>."""
    for nib in range(16):
        w = (nib * 0x00204081) & 0x01010101
        assert [(w >> (8 * q)) & 0xFF for q in range(4)] == [(nib >> q) & 1 for q in range(4)]
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    ea, eb = np.unpackbits(a, axis=1).astype(np.int64), np.unpackbits(b, axis=1).astype(np.int64)
    ham = (ea[:, None, :] != eb[None, :, :]).sum(2)
    assert np.array_equal(ham, ea.sum(1)[:, None] + eb.sum(1)[None, :] - 2 * (ea @ eb.T))
    # packed column keys: the minimum of distance << 16 | index is the smallest distance, then the smallest index
    key = (ham << 16) | np.arange(300)[:, None]
    win = key.min(0)
    assert np.array_equal(win >> 16, ham.min(0)) and np.array_equal(win & 0xFFFF, ham.argmin(0))


def test_reciprocal_multiply_index_splits():
    # (q * 993) >> 16 == q // 66, (q * 3450) >> 16 == q // 19, (e * 6554) >> 16 == e // 10 on the ranges the kernels use
    q = np.arange(0, 32768, dtype=np.int64)
    assert np.array_equal((q * 993) >> 16, q // 66)
    q = np.arange(0, 4681, dtype=np.int64)
    assert np.array_equal((q * 3450) >> 16, q // 19)
    e = np.arange(0, 16384, dtype=np.int64)
    assert np.array_equal((e * 6554) >> 16, e // 10)
    # __umulhi(t, ceil(2^32 / d)) == t // d for t * d < 2^32 (depth-edge tile list; FAST cell index uses the 2^20 form)
    rng = np.random.default_rng(4)
    for d in [2, 3, 20, 23, 40, 60, 600, 920, 2040, 4080]:
        magic = ((1 << 32) + d - 1) // d
        t = np.unique(np.concatenate([np.arange(0, 5000), rng.integers(0, (1 << 32) // d, 200000), [(1 << 32) // d - 1]])).astype(np.uint64)
        assert np.array_equal((t * np.uint64(magic)) >> np.uint64(32), t // np.uint64(d)), d
    for ncols in range(1, 80):
        inv = (1 << 20) // ncols + 1
        ci = np.arange(0, min((1 << 20) // ncols, 8192), dtype=np.int64)
        assert np.array_equal((ci * inv) >> 20, ci // ncols), ncols
