"""CPU checks of the kernel bodies (csrc/kernels.cuh, portable arithmetic path) against the oracle, through the
test-only host emulation.  These run in the GPU-less container; the same cases run on the GPU in test_gpu_parity.py."""
import ctypes
import json
import os
import random

import pytest

from oracle import ed25519 as E
from oracle import prng

P = E.P


def fe_op(lib, op, a, b=0):
    out = ctypes.create_string_buffer(32)
    lib.hostemu_fe_op(op, a.to_bytes(32, "little"), b.to_bytes(32, "little"), out)
    return int.from_bytes(out.raw, "little")


EDGE = [0, 1, 2, 19, 37, 38, P - 1, P, P + 1, 2**255 - 1, 2**255, 2**256 - 1, 2**256 - 38, 2**256 - 37, 2**256 - 39,
        2**29 - 1, 2**29, (2**256 - 1) // 3, int("1fffffff" * 8, 16)]


FE_OPS = {0: lambda x, y: x + y, 1: lambda x, y: x - y, 2: lambda x, y: x * y, 4: lambda x, y: x,
          5: lambda x, y: x * x, 6: lambda x, y: (x + y) * (x + 2 * y), 7: lambda x, y: (x - y) * 2 * (x + y),
          8: lambda x, y: (x + y) ** 2, 9: lambda x, y: (2 * x + y) - 3 * y, 10: lambda x, y: -x}


def test_field_ops(hostemu):
    """Every op returns the canonical representative, so equality is exact (not just mod p)."""
    rnd = random.Random(1)
    vals = EDGE + [rnd.getrandbits(256) for _ in range(60)]
    for a in vals:
        for b in vals[:25]:
            for op, fn in FE_OPS.items():
                assert fe_op(hostemu, op, a, b) == fn(a, b) % P, (op, a, b)
        if a % P:
            assert fe_op(hostemu, 3, a) == pow(a, -1, P)


def run_msm(lib, pts, scs, c=0, sort=1, r=3):
    n = len(scs)
    A = b"".join(E.point_to_bytes(p) for p in pts[:n])
    S = b"".join(E.scalar_to_bytes(s) for s in scs)
    out = ctypes.create_string_buffer(64)
    err = lib.hostemu_msm(A, S, n, c, sort, r, out, None)
    assert err == 0
    return E.point_from_bytes(out.raw)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 64, 300])
def test_msm_known_dlog(hostemu, known_points, n):
    dl, pts = known_points
    scs = [prng.scalar(0x5EED, i) for i in range(n)]
    exp = E.msm_known_dlog(scs, dl[:n])
    for c in ([0, 2, 3, 5, 8, 13, 16] if n <= 64 else [0, 7]):
        for r in ([1, 3, 4] if n <= 17 else [3]):
            for sort in (0, 1):
                assert run_msm(hostemu, pts, scs, c, sort, r) == exp, (n, c, r, sort)
    # the segmented accumulate kernel forced on the plain path (several segment lengths: straddling and long buckets)
    try:
        for seg_len in (0, 1, 3, 8, 50):
            hostemu.hostemu_set_seg(2, seg_len)
            for c in ([0, 3, 8, 16] if n <= 64 else [0, 7]):
                assert run_msm(hostemu, pts, scs, c, 1, 3) == exp, (n, c, seg_len)
    finally:
        hostemu.hostemu_set_seg(1, 0)


def test_msm_matches_reference_algorithm(hostemu, known_points):
    _, pts = known_points
    scs = [prng.scalar(7, i) for i in range(17)]
    assert run_msm(hostemu, pts, scs) == E.msm_naive(scs, pts)
    edge = [0, 1, E.L - 1, 2, E.L - 2, 2**252, 2**128, 2**16 - 1, 2**16, 2**15, 2**15 + 1]
    for c in (0, 16, 4):
        assert run_msm(hostemu, pts, edge, c) == E.msm_naive(edge, pts)
    pp = [pts[0]] * 5 + [E.IDENTITY] * 3 + [E.affine_neg(pts[0])] * 2
    sc = [5, 7, 11, 13, 17, 3, 4, 5, 9, 1]
    assert run_msm(hostemu, pp, sc, 4) == E.msm_naive(sc, pp)


def test_msm_skewed_scalars_long_buckets(hostemu):
    """Boolean / constant / top-window-only scalars force buckets far above the per-thread cap: exercises the
    overflow-task path (KOverflow / KCombine)."""
    n = 700
    base = [E.scalar_mul(E.B, k + 1) for k in range(40)]
    pts = [base[i % 40] for i in range(n)]
    dl = [(i % 40) + 1 for i in range(n)]
    rnd = random.Random(3)
    for scs in ([rnd.randrange(2) for _ in range(n)], [7] * n, [rnd.randrange(1, 4) * (1 << 240) for _ in range(n)]):
        for c in (0, 4, 13):
            assert run_msm(hostemu, pts, scs, c) == E.msm_known_dlog(scs, dl), c
        try:  # the same through the segmented kernel: one bucket spanning hundreds of segments (KSegLongFix)
            for seg_len in (0, 2):
                hostemu.hostemu_set_seg(2, seg_len)
                assert run_msm(hostemu, pts, scs, 4) == E.msm_known_dlog(scs, dl), seg_len
        finally:
            hostemu.hostemu_set_seg(1, 0)


def test_msm_ext_pedersen_form(hostemu, known_points):
    """vector_commitment shape: n generators from one vector + blinding base(s) from another (pivot.py:143-144)."""
    dl, pts = known_points
    hostemu.hostemu_msm_ext.restype = ctypes.c_uint32
    for n, nx in ((0, 1), (5, 1), (64, 2), (130, 1)):
        sc = [prng.scalar(0xE0, i) for i in range(n + nx)]
        main, extra = pts[:n], pts[200:200 + nx]
        out = ctypes.create_string_buffer(64)
        err = hostemu.hostemu_msm_ext(b"".join(E.point_to_bytes(p) for p in main), n,
                                      b"".join(E.point_to_bytes(p) for p in extra), nx,
                                      b"".join(E.scalar_to_bytes(s) for s in sc), 0, out)
        assert err == 0
        assert E.point_from_bytes(out.raw) == E.msm_known_dlog(sc, dl[:n] + dl[200:200 + nx])
    # equals the oracle's vector_commitment
    x, gamma = [prng.scalar(0xE1, i) for i in range(7)], prng.scalar(0xE2, 0)
    out = ctypes.create_string_buffer(64)
    hostemu.hostemu_msm_ext(b"".join(E.point_to_bytes(p) for p in pts[:7]), 7, E.point_to_bytes(E.B), 1,
                            b"".join(E.scalar_to_bytes(s) for s in x + [gamma]), 0, out)
    assert E.point_from_bytes(out.raw) == E.vector_commitment(x, gamma, pts[:7], E.B)


def test_msm_rejects_bad_points(hostemu):
    bad = (E.BX, (E.BY + 1) % P)
    out = ctypes.create_string_buffer(64)
    err = hostemu.hostemu_msm(E.point_to_bytes(bad), E.scalar_to_bytes(1), 1, 0, 1, 3, out, None)
    assert err & 2
    noncanon = int(E.BX + P).to_bytes(32, "little") + int(E.BY).to_bytes(32, "little")
    err = hostemu.hostemu_msm(noncanon, E.scalar_to_bytes(1), 1, 0, 1, 3, out, None)
    assert err & 1


def test_golden_fixture(hostemu):
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ed25519_msm_fold.json")))
    pts = [tuple(int(v, 16) for v in p) for p in g["points"]]
    for case in g["msm"]:
        sc = [int(s, 16) * (-1 if neg else 1) for s, neg in case["scalars"]]
        assert list(run_msm(hostemu, pts, sc)) == [int(v, 16) for v in case["expect"]], case["name"]
    f = g["fold"]
    A = b"".join(E.point_to_bytes(p) for p in pts[: f["n"]])
    out = ctypes.create_string_buffer(64 * (f["n"] // 2))
    hostemu.hostemu_fold(A, f["n"], E.scalar_to_bytes(int(f["c"], 16)), out)
    got = [E.point_from_bytes(out.raw[64 * i: 64 * i + 64]) for i in range(f["n"] // 2)]
    assert [[hex(x), hex(y)] for x, y in got] == f["expect"]


def test_fold(hostemu, known_points):
    _, pts = known_points
    A = b"".join(E.point_to_bytes(p) for p in pts[:32])
    out = ctypes.create_string_buffer(64 * 16)
    c = prng.scalar(99, 0)
    hostemu.hostemu_fold(A, 32, E.scalar_to_bytes(c), out)
    assert [E.point_from_bytes(out.raw[64 * i: 64 * i + 64]) for i in range(16)] == E.fold(pts[:32], c)
    for c in [0, 1, 2, 3, E.L - 1]:
        hostemu.hostemu_fold(A, 4, E.scalar_to_bytes(c), out)
        assert [E.point_from_bytes(out.raw[64 * i: 64 * i + 64]) for i in range(2)] == E.fold(pts[:4], c), c


def test_fixed_base_and_synth(hostemu, known_points):
    dl, pts = known_points
    out = ctypes.create_string_buffer(64 * 300)
    hostemu.hostemu_fixed_base(None, ctypes.c_uint64(0x5EEE), 300, out)
    assert [E.point_from_bytes(out.raw[64 * i: 64 * i + 64]) for i in range(300)] == pts
    sc = [0, 1, E.L - 1, 8, 2**252]
    hostemu.hostemu_fixed_base(b"".join(E.scalar_to_bytes(s) for s in sc), ctypes.c_uint64(0), len(sc), out)
    assert [E.point_from_bytes(out.raw[64 * i: 64 * i + 64]) for i in range(len(sc))] == [E.scalar_mul(E.B, s) for s in sc]
    out = ctypes.create_string_buffer(32 * 300)
    hostemu.hostemu_synth_scalars(ctypes.c_uint64(0x5EED), 300, out)
    assert [int.from_bytes(out.raw[32 * i: 32 * i + 32], "little") for i in range(300)] == \
        [prng.scalar(0x5EED, i) for i in range(300)]


def test_numpy_synth_matches_oracle_prng():
    from verifiable_mpc_b200 import synth

    a = synth.scalars_ed25519(0x5EED, 2048)
    for i in [0, 1, 2, 1000, 2047]:
        assert int.from_bytes(a[i].tobytes(), "little") == prng.scalar(0x5EED, i)
    b = synth.scalars_ed25519(0x5EED, 8, start=1000)
    assert (b[0] == a[1000]).all()


def test_point_text_matches_python_repr(hostemu, known_points):
    """KPointText: the decimal text of the Fiat-Shamir pre-image, byte for byte what Python prints."""
    _, pts = known_points
    cases = pts[:20] + [E.IDENTITY, (0, 1), (1, 0), (10**9, 10**18 - 1), (10**9 - 1, 10**9), (P - 1, 2**255 - 20)]
    n = len(cases)
    slots = ctypes.create_string_buffer(176 * n)
    lens = (ctypes.c_uint32 * n)()
    hostemu.hostemu_point_text(b"".join(E.point_to_bytes(p) for p in cases), n, slots, lens)
    got = "".join(slots.raw[176 * i: 176 * i + lens[i]].decode() for i in range(n))
    assert got == ", ".join(f"[{x}, {y}, 1]" for x, y in cases)


# ---- scalar vectors modulo the group order (sc25519.cuh): witness / linear-form halving, cross terms, text
def _aligned_scalars(vals):
    """n x 32 bytes, 16-byte aligned (the kernels use 128-bit loads)."""
    import numpy as np
    arr = np.zeros(len(vals) * 8 + 4, dtype=np.uint32)
    off = (-arr.ctypes.data % 16) // 4
    view = arr[off:off + len(vals) * 8]
    view[:] = np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint32)
    return view


def _edge_scalars(n, seed):
    import random
    rng = random.Random(seed)
    L = E.L
    vals = [0, 1, L - 1, L >> 1, (L >> 1) + 1, 2**252, 2**252 - 1, 10**9, 10**9 - 1, 2**32 - 1, 2**32]
    vals += [rng.randrange(L) for _ in range(n - len(vals))]
    return vals[:n]


@pytest.mark.parametrize("mode", [0, 1])
def test_scalars_fold(hostemu, mode):
    """KScalarFold against Python ints: z' = z_L + c z_R (mode 0) and L' = c L_L + L_R (mode 1), edge values included."""
    L = E.L
    for half, c in ((1, 0), (7, 1), (16, L - 1), (33, 0x1234567890ABCDEF << 180)):
        vals = _edge_scalars(2 * half, half) if half > 5 else [L - 1, 5][:2 * half] + [3] * max(0, 2 * half - 2)
        v = _aligned_scalars(vals)
        hostemu.hostemu_scalars_fold(v.ctypes.data_as(ctypes.c_void_p), half, (c % L).to_bytes(32, "little"), mode)
        got = [int.from_bytes(v[8 * i:8 * i + 8].tobytes(), "little") for i in range(half)]
        lo, hi = vals[:half], vals[half:2 * half]
        exp = [(a + c * b) % L for a, b in zip(lo, hi)] if mode == 0 else [(c * a + b) % L for a, b in zip(lo, hi)]
        assert got == exp


def test_scalars_axpy(hostemu):
    L = E.L
    n = 23
    d_vals, s_vals = _edge_scalars(n, 1), list(reversed(_edge_scalars(n, 2)))
    for mode, c in ((0, 5), (1, L - 1), (2, 2**251 + 12345), (2, 0)):
        d, sv = _aligned_scalars(d_vals), _aligned_scalars(s_vals)
        hostemu.hostemu_scalars_axpy(d.ctypes.data_as(ctypes.c_void_p), sv.ctypes.data_as(ctypes.c_void_p), n,
                                     c.to_bytes(32, "little"), mode)
        got = [int.from_bytes(d[8 * i:8 * i + 8].tobytes(), "little") for i in range(n)]
        exp = [(a + c * b) % L for a, b in zip(d_vals, s_vals)] if mode == 0 else \
            [(c * a + b) % L for a, b in zip(d_vals, s_vals)] if mode == 1 else [c * a % L for a in d_vals]
        assert got == exp


def test_scalars_dot(hostemu):
    L = E.L
    for n in (0, 1, 2, 63, 64, 65, 255, 256, 257, 300, 5000, 66000):
        a_vals, b_vals = _edge_scalars(max(n, 11), n)[:n], list(reversed(_edge_scalars(max(n, 11), n + 1)))[:n]
        a, b = _aligned_scalars(a_vals or [0]), _aligned_scalars(b_vals or [0])
        out = ctypes.create_string_buffer(32)
        hostemu.hostemu_scalars_dot(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), n, out)
        assert int.from_bytes(out.raw, "little") == sum(x * y for x, y in zip(a_vals, b_vals)) % L


@pytest.mark.parametrize("signed", [0, 1])
def test_scalar_text_matches_python_repr(hostemu, signed):
    """KScalarText: the coefficient text of L_tilde exactly as MPyC prints (signed) field elements."""
    L = E.L
    vals = _edge_scalars(40, 5)
    n = len(vals)
    v = _aligned_scalars(vals)
    slots = ctypes.create_string_buffer(96 * n)
    lens = (ctypes.c_uint32 * n)()
    hostemu.hostemu_scalar_text(v.ctypes.data_as(ctypes.c_void_p), n, signed, slots, lens)
    got = "".join(slots.raw[96 * i: 96 * i + lens[i]].decode() for i in range(n))
    exp = ", ".join(str(x - L if signed and x > (L >> 1) else x) for x in vals)
    assert got == exp


def test_bench_window_mirror_matches_library(hostemu):
    """bench.py reports the window the library will choose (for the roofline arithmetic): the Python mirror must agree
    with csrc/pipeline.cuh:choose_window at every size the bench or the sweep can be asked for."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    sizes = [1, 2, 100] + [1 << k for k in range(4, 27)] + [49151, 49152, 3 << 15, (1 << 20) + 5]
    for n in sizes:
        assert bench._choose_window(n) == hostemu.hostemu_choose_window(ctypes.c_uint64(n)), n


def run_msm_pre(lib, pts, off, n_main, extra, scs, table_bits, sets):
    A = b"".join(E.point_to_bytes(p) for p in pts)
    X = b"".join(E.point_to_bytes(p) for p in extra)
    S = b"".join(E.scalar_to_bytes(s) for s in scs)
    out = ctypes.create_string_buffer(64)
    rc = lib.hostemu_msm_pre(A, len(pts), off, n_main, X, len(extra), S, table_bits, sets, out)
    assert rc == 0
    return E.point_from_bytes(out.raw)


@pytest.mark.parametrize("table_bits", [8, 13, 16])
def test_msm_over_precomputed_bases(hostemu, known_points, table_bits):
    """Tables of 2^(c*w) * P_i (KPrecompute): windows share bucket sets, no doublings; any number of sets, any
    sub-range of the vector, extra terms with their own table, edge scalars, repeated / identity / negated bases."""
    dl, pts = known_points
    W = (253 + table_bits) // table_bits
    n = 40
    scs = [prng.scalar(0x5EED, i) for i in range(n)]
    scs[:6] = [0, 1, E.L - 1, 2**252, 2**(table_bits - 1), 2**table_bits - 1]
    want = E.msm_known_dlog(scs, dl[:n])
    try:
        for mode, seg_len in ((1, 0), (0, 0), (2, 5), (2, 64)):  # by geometry / one thread per bucket / segments
            hostemu.hostemu_set_seg(mode, seg_len)
            for sets in sorted({0, 1, 2, 3, 5, W - 1, W}):
                assert run_msm_pre(hostemu, pts[:n], 0, n, [], scs, table_bits, sets) == want, (mode, seg_len, sets)
    finally:
        hostemu.hostemu_set_seg(1, 0)
    # sub-range of a longer vector + two extra terms (the h / k of a commitment)
    extra = [pts[200], E.affine_neg(pts[201])]
    sc2 = scs[:20] + [prng.scalar(9, 0), prng.scalar(9, 1)]
    want2 = E.msm_known_dlog(sc2, dl[7:27] + [dl[200], E.L - dl[201]])
    for sets in (0, 1, 4):
        assert run_msm_pre(hostemu, pts[:60], 7, 20, extra, sc2, table_bits, sets) == want2, sets
    pp = [pts[0]] * 5 + [E.IDENTITY] * 3 + [E.affine_neg(pts[0])] * 2
    sc = [5, 7, 11, 13, 17, 3, 4, 5, 9, 1]
    assert run_msm_pre(hostemu, pp, 0, len(pp), [], sc, table_bits, 2) == E.msm_naive(sc, pp)
    assert run_msm_pre(hostemu, pts[:3], 0, 0, [], [], table_bits, 0) == E.IDENTITY


def test_block_privatised_sort_gives_the_same_sums(hostemu, known_points):
    """The block-privatised counting sort (digits recoded once into 16-bit codes, per-(set, chunk) histograms, chunk
    prefix, per-block cursors: vmsm_bsort_* on the device, restated as loops in the emulation around the shared KRecode
    body) feeds the same accumulate kernels: plain path, shared bucket sets over tables, segments, skewed scalars."""
    dl, pts = known_points
    n = 203  # not a multiple of 8: ragged last vector of codes, ragged last chunk
    scs = [prng.scalar(0xB5, i) for i in range(n)]
    scs[:7] = [0, 1, E.L - 1, 2**252, 2**15, 2**15 + 1, 2**16 - 1]
    want = E.msm_known_dlog(scs, dl[:n])
    try:
        for chunks in (1, 3, 8, 300):  # 300 > n / 8: most chunks are empty
            hostemu.hostemu_set_block_sort(chunks)
            for c in (0, 4, 11, 16):
                assert run_msm(hostemu, pts[:n], scs, c, 1, 3) == want, (chunks, c)
            for mode, seg_len in ((1, 0), (0, 0), (2, 5)):
                hostemu.hostemu_set_seg(mode, seg_len)
                for sets in (0, 1, 3):
                    assert run_msm_pre(hostemu, pts[:n], 0, n, [], scs, 8, sets) == want, (chunks, mode, seg_len, sets)
            hostemu.hostemu_set_seg(1, 0)
        rnd = random.Random(5)
        bits = [rnd.randrange(2) for _ in range(n)]
        hostemu.hostemu_set_block_sort(4)
        assert run_msm(hostemu, pts[:n], bits, 13) == E.msm_known_dlog(bits, dl[:n])
        assert run_msm_pre(hostemu, pts[:n], 0, n, [], bits, 8, 1) == E.msm_known_dlog(bits, dl[:n])
    finally:
        hostemu.hostemu_set_seg(1, 0)
        hostemu.hostemu_set_block_sort(0)


def test_msm_over_precomputed_bases_long_buckets(hostemu):
    """Boolean / constant scalars over shared bucket sets: every window's entries pile into a few buckets (overflow
    tasks across windows)."""
    n = 600
    base = [E.scalar_mul(E.B, k + 1) for k in range(30)]
    pts = [base[i % 30] for i in range(n)]
    rnd = random.Random(4)
    for sc in ([rnd.randrange(2) for _ in range(n)], [7] * n, [rnd.randrange(1, 4) << 240 for _ in range(n)]):
        want = E.scalar_mul(E.B, sum(s * (i % 30 + 1) for i, s in enumerate(sc)) % E.L)
        try:
            for mode, seg_len in ((1, 0), (0, 0), (2, 2)):  # seg_len 2: buckets spanning > 32 segments (long fix-up)
                hostemu.hostemu_set_seg(mode, seg_len)
                for sets in (1, 2):
                    assert run_msm_pre(hostemu, pts, 0, n, [], sc, 8, sets) == want, (mode, seg_len, sets)
        finally:
            hostemu.hostemu_set_seg(1, 0)
