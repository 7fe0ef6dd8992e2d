"""CPU checks of the BN256 kernel bodies (csrc/fbn256.cuh, bn256.cuh, kernels_w.cuh) against oracle/bn256.py through
the host emulation: Montgomery field, G1 and G2 group law incl. every exceptional case, fixed-base comb, MSM."""
import ctypes
import random

import pytest

from oracle import bn256 as B
from oracle import prng

P = B.P


def fop(lib, op, a, b=0):
    out = ctypes.create_string_buffer(32)
    lib.hostemu_fbn_op(op, a.to_bytes(32, "little"), b.to_bytes(32, "little"), out)
    return int.from_bytes(out.raw, "little")


def test_oracle_constants():
    assert B.on_curve(B.FP, B.G1) and B.on_curve(B.FP2, B.G2)
    assert B.scalar_mul(B.FP, B.G1, B.N) is None and B.scalar_mul(B.FP2, B.G2, B.N) is None
    assert B.P % 4 == 3 and B.P.bit_length() == 256 and B.N.bit_length() == 256
    for F, G in ((B.FP, B.G1), (B.FP2, B.G2)):
        a, b = B.scalar_mul(F, G, 12345), B.scalar_mul(F, G, 67890)
        assert B.affine_add(F, a, b) == B.scalar_mul(F, G, 12345 + 67890)
        assert B.affine_add(F, a, a) == B.scalar_mul(F, G, 2 * 12345)
        assert B.affine_add(F, a, B.affine_neg(F, a)) is None
        assert B.scalar_mul(F, G, -5) == B.affine_neg(F, B.scalar_mul(F, G, 5))


def test_montgomery_field(hostemu):
    rnd = random.Random(1)
    vals = [v % P for v in [0, 1, 2, P - 1, P - 2, (P - 1) // 2, 2**255, 2**255 + 12345]] + \
        [rnd.randrange(P) for _ in range(60)]
    for a in vals:
        for b in vals[:20]:
            assert fop(hostemu, 0, a, b) == (a + b) % P
            assert fop(hostemu, 1, a, b) == (a - b) % P
            assert fop(hostemu, 2, a, b) == a * b % P
        assert fop(hostemu, 4, a) == (-a) % P
        if a:
            assert fop(hostemu, 3, a) == pow(a, -1, P)


@pytest.mark.parametrize("g2", [0, 1])
def test_group_fixed_base_and_msm(hostemu, g2):
    lib = hostemu
    lib.hostemu_bn_msm.restype = ctypes.c_uint32
    F, G = (B.FP2, B.G2) if g2 else (B.FP, B.G1)
    sz, n = (128 if g2 else 64), 40
    dl = [prng.scalar_bn(0x700 + g2, i) for i in range(n)]
    out = ctypes.create_string_buffer(sz * n)
    lib.hostemu_bn_fixed_base(g2, b"".join(B.fp_to_bytes(d) for d in dl), ctypes.c_uint64(0), n, out)
    pts = [B.point_from_bytes(F, out.raw[sz * i: sz * i + sz]) for i in range(n)]
    assert pts[:6] == [B.scalar_mul(F, G, d) for d in dl[:6]]
    out2 = ctypes.create_string_buffer(sz * 4)
    lib.hostemu_bn_fixed_base(g2, None, ctypes.c_uint64(0x700 + g2), 4, out2)
    assert [B.point_from_bytes(F, out2.raw[sz * i: sz * i + sz]) for i in range(4)] == pts[:4]

    def msm(points, scalars, c=0):
        o = ctypes.create_string_buffer(sz)
        err = lib.hostemu_bn_msm(g2, b"".join(B.point_to_bytes(F, p) for p in points),
                                 b"".join(B.fp_to_bytes(s % B.N) for s in scalars), len(scalars), c, o)
        return err, B.point_from_bytes(F, o.raw)

    for nn, c in ((0, 0), (1, 0), (3, 4), (17, 0), (40, 5), (40, 0), (40, 11)):
        sc = [prng.scalar_bn(0x800, i) for i in range(nn)]
        err, got = msm(pts[:nn], sc, c)
        assert err == 0 and got == B.msm_known_dlog(F, sc, dl[:nn]), (nn, c)
    # exceptional cases of the incomplete formulas: repeated bases (doubling), P and -P (cancellation), identity base
    pp = [pts[0]] * 4 + [B.affine_neg(F, pts[0])] * 2 + [None] + [pts[1]]
    sc = [5, 5, 7, 1, 3, 9, 1234, B.N - 1]
    assert msm(pp, sc, 3) == (0, B.msm_naive(F, sc, pp))
    assert msm(pp, [1, 1, 1, 1, 2, 2, 0, 0], 4) == (0, None)
    # matches the reference's algorithm (per-term double-and-add + apply_to_list tree)
    sc = [prng.scalar_bn(0x801, i) for i in range(9)]
    assert msm(pts[:9], sc)[1] == B.msm_naive(F, sc, pts[:9])
    # validation
    assert msm([(pts[0][0], pts[1][1])], [1])[0] & 2


@pytest.mark.parametrize("g2", [0, 1])
def test_msm_over_key_tables(hostemu, g2):
    """KPrecomputeW tables (2^(c*w) * P_i): no doublings in the MSM, every window its own bucket set; sub-range of the
    vector, extra terms with their own table, the exceptional cases (repeated / negated / identity bases), and the
    host-side normalisation of the Jacobian result."""
    lib = hostemu
    F, G = (B.FP2, B.G2) if g2 else (B.FP, B.G1)
    sz, n = (128 if g2 else 64), 24
    dl = [prng.scalar_bn(0x710 + g2, i) for i in range(n)]
    pts = [B.scalar_mul(F, G, d) for d in dl]

    def msm_pre(points, off, n_main, extra, scalars, c):
        o = ctypes.create_string_buffer(sz)
        rc = lib.hostemu_bn_msm_pre(g2, b"".join(B.point_to_bytes(F, p) for p in points), len(points), off, n_main,
                                    b"".join(B.point_to_bytes(F, p) for p in extra), len(extra),
                                    b"".join(B.fp_to_bytes(s % B.N) for s in scalars), c, o)
        assert rc == 0
        return B.point_from_bytes(F, o.raw)

    sc = [prng.scalar_bn(0x810, i) for i in range(n)]
    sc[:4] = [0, 1, B.N - 1, 2**255]
    for c in (8, 13):
        assert msm_pre(pts, 0, n, [], sc, c) == B.msm_known_dlog(F, sc, dl), c
    extra = [pts[3], B.affine_neg(F, pts[5])]
    sc2 = sc[:10] + [7, 9]
    assert msm_pre(pts, 4, 10, extra, sc2, 13) == B.msm_known_dlog(F, sc2, dl[4:14] + [dl[3], B.N - dl[5]])
    pp = [pts[0]] * 4 + [B.affine_neg(F, pts[0])] * 2 + [None] + [pts[1]]
    s3 = [5, 5, 7, 1, 3, 9, 1234, B.N - 1]
    assert msm_pre(pp, 0, len(pp), [], s3, 8) == B.msm_naive(F, s3, pp)
    assert msm_pre(pp, 0, len(pp), [], [1, 1, 1, 1, 2, 2, 0, 0], 8) is None


@pytest.mark.parametrize("g2", [0, 1])
def test_msm_over_key_tables_shared_sets_and_segments(hostemu, g2):
    """Windows sharing S bucket sets over the key tables (CSR entries carry which window: BaseRefW) summed by one thread
    per bucket (+ long-bucket tasks) or by equal segments (KAccumulateSegW / KSegFixupW / KSegLongFixW): every
    combination gives the known-dlog result, also for skewed scalars whose entries pile into one bucket."""
    import random

    lib = hostemu
    F, G = (B.FP2, B.G2) if g2 else (B.FP, B.G1)
    sz, n = (128 if g2 else 64), 40 if g2 else 260  # >= 256 terms on G1: the segments are the default there
    base = [B.scalar_mul(F, G, k + 1) for k in range(10)]
    pts = [base[i % 10] for i in range(n)]
    dl = [i % 10 + 1 for i in range(n)]

    def msm_pre(scalars, c):
        o = ctypes.create_string_buffer(sz)
        rc = lib.hostemu_bn_msm_pre(g2, b"".join(B.point_to_bytes(F, p) for p in pts), n, 0, n, b"", 0,
                                    b"".join(B.fp_to_bytes(s % B.N) for s in scalars), c, o)
        assert rc == 0
        return B.point_from_bytes(F, o.raw)

    rnd = random.Random(11 + g2)
    cases = [[prng.scalar_bn(0x820, i) for i in range(n)], [rnd.randrange(2) for _ in range(n)], [7] * n]
    try:
        for mode, seg_len in ((1, 0), (0, 0), (2, 3)):  # n >= 256 on G1: segments by default; 40 terms on G2: forced
            hostemu.hostemu_set_seg(mode, seg_len)
            for sets in ((0, 1, 3) if g2 else (0, 1, 5, 32)):
                hostemu.hostemu_set_bn_sets(sets)
                for sc in cases[:2] if g2 or sets else cases:
                    assert msm_pre(sc, 8) == B.msm_known_dlog(F, sc, dl), (mode, seg_len, sets)
    finally:
        hostemu.hostemu_set_seg(1, 0)
        hostemu.hostemu_set_bn_sets(0)


@pytest.mark.parametrize("g2", [0, 1])
def test_plain_path_with_forced_segments(hostemu, g2):
    """VMSM_OPT_SEG_MODE 2 on the plain path (own bucket set per window, Horner chain kept): same sums."""
    lib = hostemu
    lib.hostemu_bn_msm.restype = ctypes.c_uint32
    F, G = (B.FP2, B.G2) if g2 else (B.FP, B.G1)
    sz, n = (128 if g2 else 64), 300
    base = [B.scalar_mul(F, G, k + 1) for k in range(8)]
    pts = [base[i % 8] for i in range(n)]
    dl = [i % 8 + 1 for i in range(n)]
    sc = [prng.scalar_bn(0x840, i) for i in range(n)]
    sc[:3] = [0, B.N - 1, 1]
    try:
        for seg_len in (0, 5):
            hostemu.hostemu_set_seg(2, seg_len)
            for c in ((0, 9) if not g2 else (9,)):
                o = ctypes.create_string_buffer(sz)
                err = lib.hostemu_bn_msm(g2, b"".join(B.point_to_bytes(F, p) for p in pts),
                                         b"".join(B.fp_to_bytes(s) for s in sc), n, c, o)
                assert err == 0
                assert B.point_from_bytes(F, o.raw) == B.msm_known_dlog(F, sc, dl), (seg_len, c)
    finally:
        hostemu.hostemu_set_seg(1, 0)
