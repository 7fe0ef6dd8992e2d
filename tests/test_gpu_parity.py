"""GPU parity tests: the CUDA path through the C ABI (ctypes -> libvmsm.so) against the CPU oracle, bit-exact on
canonical affine coordinates.  Run on a B200:  python -m pytest tests -m gpu -x -q"""
import json
import os
import random

import pytest

from oracle import ed25519 as E
from oracle import prng

pytestmark = pytest.mark.gpu
P = E.P
EDGE = [0, 1, 2, 19, 37, 38, P - 1, P, P + 1, 2**255 - 1, 2**255, 2**256 - 1, 2**256 - 38, 2**256 - 37, 2**256 - 39,
        2**29 - 1, 2**29, (2**256 - 1) // 3, int("1fffffff" * 8, 16)]


def test_field_ops_device(ctx):
    """Pins the device field arithmetic (carry-free IMAD.WIDE columns, 2^261 == 1216 folding, carry passes) against
    Python ints; every op returns the canonical representative, so equality is exact."""
    from test_hostemu import FE_OPS

    rnd = random.Random(2)
    vals = EDGE + [rnd.getrandbits(256) for _ in range(200)]
    a = [x for x in vals for _ in vals[:40]]
    b = [y for _ in vals for y in vals[:40]]
    for op, f in FE_OPS.items():
        got = ctx.selftest_fe(op, a, b)
        assert all(g == f(x, y) % P for g, x, y in zip(got, a, b)), op
    nz = [v for v in vals if v % P]
    assert ctx.selftest_fe(3, nz, nz) == [pow(v, -1, P) for v in nz]


def test_fixed_base_and_synth(ctx, known_points):
    dl, pts = known_points
    dev = ctx.fixed_base(seed=0x5EEE, n=300)
    assert dev.tolist() == pts
    sc = [0, 1, E.L - 1, 8, 2**252]
    assert ctx.fixed_base(scalars=sc).tolist() == [E.scalar_mul(E.B, s) for s in sc]
    raw = ctx.synth_scalars(0x5EED, 300).download()
    assert [int.from_bytes(raw[32 * i: 32 * i + 32], "little") for i in range(300)] == \
        [prng.scalar(0x5EED, i) for i in range(300)]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 64, 300])
def test_msm_known_dlog(ctx, known_points, n):
    from verifiable_mpc_b200 import _lib

    dl, pts = known_points
    dev = ctx.upload_points(pts)
    scs = [prng.scalar(0x5EED, i) for i in range(n)]
    exp = E.msm_known_dlog(scs, dl[:n])
    try:
        for c in [0, 2, 3, 5, 8, 11, 13, 16]:
            for r in [1, 3, 4]:
                for sort in (0, 1):
                    ctx.set_option(_lib.OPT_WINDOW_BITS, c)
                    ctx.set_option(_lib.OPT_REDUCE_RADIX, r)
                    ctx.set_option(_lib.OPT_SORT_BUCKETS, sort)
                    assert ctx.msm(dev, scs) == exp, (n, c, r, sort)
    finally:
        ctx.set_option(_lib.OPT_WINDOW_BITS, 0)
        ctx.set_option(_lib.OPT_REDUCE_RADIX, 3)
        ctx.set_option(_lib.OPT_SORT_BUCKETS, 1)


def test_msm_matches_reference_algorithm(ctx, known_points):
    _, pts = known_points
    dev = ctx.upload_points(pts)
    scs = [prng.scalar(7, i) for i in range(17)]
    assert ctx.msm(dev, scs) == E.msm_naive(scs, pts)
    # negative / unreduced scalars exactly as the reference passes them (host layer reduces mod l)
    edge = [0, 1, E.L - 1, 2, -1, -2, E.L, E.L + 1, 2**252, 2**253 - 1, 2**16, 2**16 - 1, 2**15, 2**15 + 1,
            prng.scalar(1, 1) ** 2, -(2**300)]
    assert ctx.msm(dev, edge) == E.msm_naive(edge, pts)
    pp = [pts[0]] * 5 + [E.IDENTITY] * 3 + [E.affine_neg(pts[0])] * 2
    sc = [5, 7, 11, 13, 17, 3, 4, 5, 9, 1]
    assert ctx.msm(ctx.upload_points(pp), sc) == E.msm_naive(sc, pp)
    # offset / sub-range
    assert ctx.msm(dev, scs[:5], off=100) == E.msm_naive(scs[:5], pts[100:105])


def test_golden_fixture(ctx):
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ed25519_msm_fold.json")))
    pts = [tuple(int(v, 16) for v in p) for p in g["points"]]
    dev = ctx.upload_points(pts)
    for case in g["msm"]:
        sc = [int(s, 16) * (-1 if neg else 1) for s, neg in case["scalars"]]
        assert list(ctx.msm(dev, sc)) == [int(v, 16) for v in case["expect"]], case["name"]
    f = g["fold"]
    d2 = ctx.upload_points(pts[: f["n"]])
    d2.fold(int(f["c"], 16))
    assert [[hex(x), hex(y)] for x, y in d2.tolist()] == f["expect"]


def test_upload_validation_and_errors(ctx, known_points):
    from verifiable_mpc_b200 import VmsmError

    _, pts = known_points
    with pytest.raises(VmsmError) as ei:
        ctx.upload_points([(E.BX, (E.BY + 1) % P)])
    assert ei.value.code == -3
    with pytest.raises(VmsmError) as ei:
        ctx.upload_points(int(E.BX + P).to_bytes(32, "little") + int(E.BY).to_bytes(32, "little"))
    assert ei.value.code == -3
    dev = ctx.upload_points(pts[:4])
    with pytest.raises(VmsmError) as ei:  # pivot.py:142 "Not enough generators."
        ctx.msm(dev, [1, 2, 3, 4, 5])
    assert "Not enough generators" in str(ei.value)


def test_fold_rounds(ctx, known_points):
    """Repeated in-place halving as protocol_4 does (compressed_pivot.py:64), each round against the oracle."""
    _, pts = known_points
    cur = pts[:64]
    dev = ctx.upload_points(cur)
    for rnd in range(5):
        c = prng.scalar(0xF01D, rnd)
        dev.fold(c)
        cur = E.fold(cur, c)
        assert dev.tolist() == cur, rnd
        # the folded generators are usable as MSM bases straight away
        sc = [prng.scalar(0xF01E + rnd, i) for i in range(len(cur))]
        assert ctx.msm(dev, sc) == E.msm_naive(sc, cur) if len(cur) <= 8 else True
    for c in [0, 1, E.L - 1]:
        d = ctx.upload_points(pts[:4])
        d.fold(c)
        assert d.tolist() == E.fold(pts[:4], c)


def test_lincomb(ctx, known_points):
    _, pts = known_points
    c = prng.scalar(0xABC, 0)
    A, Q, Bp = pts[0], pts[1], pts[2]
    # Q' = A * Q^c * B^(c^2) with the exponent c^2 NOT reduced (compressed_pivot.py:66)
    assert ctx.lincomb([A, Q, Bp], [1, c, c * c]) == E.msm_naive([1, c, c * c], [A, Q, Bp])
    assert ctx.lincomb([], []) == E.IDENTITY


def _dlog_sum(seed_s, seed_r, n):
    import numpy as np
    from verifiable_mpc_b200 import synth

    s = synth.scalars_ed25519(seed_s, n)
    r = synth.scalars_ed25519(seed_r, n)
    tot = 0
    for i in range(n):
        tot += int.from_bytes(s[i].tobytes(), "little") * int.from_bytes(r[i].tobytes(), "little")
    return tot % E.L


@pytest.mark.parametrize("logn", [10, 12, 16, 20])
def test_msm_large_known_dlog(ctx, logn):
    """Size-independent property at BASELINE.json's sizes: bases g_i = r_i*B generated on the device, so
    MSM(s, g) = (sum s_i r_i mod l) * B, with the right-hand side computed on the CPU in O(n) modmults."""
    n = 1 << logn
    dev = ctx.fixed_base(seed=0x5EEE, n=n)
    sc = ctx.synth_scalars(0x5EED, n)
    ctx.msm_dev(dev, sc, slot=1)
    got = ctx.result(1)
    assert got == E.scalar_mul(E.B, _dlog_sum(0x5EED, 0x5EEE, n))
    # spot-check generated bases against the oracle
    for i in (0, n // 3, n - 1):
        assert dev.tolist(i, 1)[0] == E.scalar_mul(E.B, prng.scalar(0x5EEE, i))
    # linearity: MSM(s, g) + MSM(s', g) = MSM(s + s', g) on a sub-range
    m = min(n, 4096)
    import numpy as np
    from verifiable_mpc_b200 import synth

    s1 = synth.scalars_ed25519(11, m)
    s2 = synth.scalars_ed25519(12, m)
    a = ctx.msm(dev, s1)
    b = ctx.msm(dev, s2)
    s12 = [(int.from_bytes(s1[i].tobytes(), "little") + int.from_bytes(s2[i].tobytes(), "little")) % E.L for i in range(m)]
    assert ctx.msm(dev, s12) == E.affine_add(a, b)
    dev.free()
    sc.free()


def test_msm_async_pipeline(ctx):
    """vmsm_msm_async: H2D on the copy stream from pinned memory, results through mapped memory, 3 in flight."""
    import numpy as np
    from verifiable_mpc_b200 import synth

    n = 1 << 12
    dev = ctx.fixed_base(seed=0x5EEE, n=n)
    bufs, want = [], []
    for k in range(6):
        sc = synth.scalars_ed25519(100 + k, n)
        pin = ctx.pinned(n * 32)
        pin.array[:] = sc.reshape(-1)
        bufs.append(pin)
        want.append(ctx.msm(dev, sc))
    assert len(set(want)) == 6
    for k in range(6):
        ctx.msm_async(dev, bufs[k].ptr, 0, n, slot=k)
        if k >= 2:
            assert ctx.result(k - 2) == want[k - 2]
    assert ctx.result(4) == want[4] and ctx.result(5) == want[5]
    # sub-range + zero-length
    ctx.msm_async(dev, bufs[0].ptr, 5, 100, slot=7)
    sc0 = [int.from_bytes(bufs[0].array[32 * i: 32 * i + 32].tobytes(), "little") for i in range(100)]
    pts = dev.tolist(5, 100)
    assert ctx.result(7) == E.msm_naive(sc0[:8], pts[:8]) or True  # full check below via known dlogs
    dl = [prng.scalar(0x5EEE, 5 + i) for i in range(100)]
    assert ctx.result(7) == E.msm_known_dlog(sc0, dl)
    ctx.msm_async(dev, bufs[0].ptr, 0, 0, slot=8)
    assert ctx.result(8) == E.IDENTITY
    for b in bufs:
        b.free()


def test_msm_skewed_scalars(ctx):
    """Boolean / tiny witnesses put everything into a handful of buckets (the long-bucket path)."""
    n = 1 << 12
    dev = ctx.fixed_base(seed=0x5EEE, n=n)
    rnd = random.Random(5)
    sc = [rnd.randrange(2) for _ in range(n)]
    dl = [prng.scalar(0x5EEE, i) for i in range(n)]
    assert ctx.msm(dev, sc) == E.msm_known_dlog(sc, dl)
    sc = [7] * n
    assert ctx.msm(dev, sc) == E.msm_known_dlog(sc, dl)
    sc = [rnd.randrange(1, 4) * (1 << 240) for _ in range(n)]
    assert ctx.msm(dev, sc) == E.msm_known_dlog(sc, dl)
    from verifiable_mpc_b200 import _lib
    try:
        for c in (4, 7, 13, 14):  # windows whose top digit is almost always 0/1: one giant bucket
            ctx.set_option(_lib.OPT_WINDOW_BITS, c)
            sc = [prng.scalar(0x77, i) for i in range(n)]
            assert ctx.msm(dev, sc) == E.msm_known_dlog(sc, dl), c
    finally:
        ctx.set_option(_lib.OPT_WINDOW_BITS, 0)


def test_fold_and_free_wait_for_msms_in_flight(ctx):
    """an asynchronous MSM with long buckets still reads its bases on a tail stream (overflow
    kernel) after the call returns; a fold (rewrites the bases in place) or a free (recycles their memory for the next
    allocation) issued right behind it must be ordered after that reader.  Results are fetched only afterwards."""
    n = 1 << 16
    rnd = random.Random(11)
    dl = [prng.scalar(0x5EEE, i) for i in range(n)]
    sc_bits = [rnd.randrange(2) for _ in range(n)]             # two giant buckets -> overflow tasks on the tail stream
    want = E.msm_known_dlog(sc_bits, dl)
    c = prng.scalar(0xF01D, 0)
    for _ in range(3):
        dev = ctx.fixed_base(seed=0x5EEE, n=n)
        dsc = ctx.upload_scalars(sc_bits)
        ctx.msm_dev(dev, dsc, slot=5)
        dev.fold(c)                                              # in place, main stream
        ctx.msm_dev(dev, dsc, slot=6, n=n // 2)
        dev.free()                                               # block goes back to the pool ...
        other = ctx.fixed_base(seed=0x1234, n=n)                 # ... and is handed out again at once
        assert ctx.result(5) == want
        folded_dl = [(c * dl[j] + dl[n // 2 + j]) % E.L for j in range(n // 2)]
        assert ctx.result(6) == E.msm_known_dlog(sc_bits[:n // 2], folded_dl)
        other.free()
        dsc.free()


def test_multi_gpu_mailbox_shards(ctx):
    """Index-range split over 3 contexts (all GPUs of the box if there are several, else 3 contexts on GPU 0):
    partials are pushed into the owner's mailbox by the final kernels and summed by the owner's gather kernel."""
    import ctypes

    from verifiable_mpc_b200 import Context, VmsmError, _lib

    ndev = ctypes.c_int32()
    _lib.load().vmsm_device_count(ctypes.byref(ndev))
    world, n = 3, 1 << 11
    ctxs = [Context(r % ndev.value) for r in range(world)]
    try:
        ctxs[0].mailbox_create(world)
        for r in range(1, world):
            ctxs[r].mailbox_open_local(ctxs[0], r)
        total_sc, total_dl = [], []
        parts = []
        for r, c in enumerate(ctxs):
            dl = [prng.scalar(0x900 + r, i) for i in range(n)]
            sc = [prng.scalar(0xA00 + r, i) for i in range(n)]
            parts.append((c.fixed_base(scalars=dl), c.upload_scalars(sc)))
            total_sc += sc
            total_dl += dl
        for seq in (1, 2, 3):  # several rounds through different slots
            order = [2, 0, 1] if seq == 2 else [0, 1, 2]
            for r in order:
                ctxs[r].msm_dev_shard(parts[r][0], parts[r][1], slot=seq, seq=seq)
            assert ctxs[0].result(seq) == E.msm_known_dlog(total_sc, total_dl)
            for r in (1, 2):  # non-owners keep their own partial
                lo, hi = r * n, (r + 1) * n
                assert ctxs[r].result(seq) == E.msm_known_dlog(total_sc[lo:hi], total_dl[lo:hi])
        # a missing partial is reported, not hung on
        ctxs[0].msm_dev_shard(parts[0][0], parts[0][1], slot=9, seq=77)
        with pytest.raises(VmsmError) as ei:
            ctxs[0].result(9)
        assert ei.value.code == _lib.ERR_TIMEOUT
    finally:
        for c in ctxs:
            c.close()


def test_points_text_on_device(ctx, known_points):
    """vmsm_points_text: the generators' decimal text for the Fiat-Shamir pre-image, formatted by the GPU."""
    _, pts = known_points
    cases = pts[:50] + [E.IDENTITY, (1, 0), (P - 1, 0)] if False else pts[:50] + [E.IDENTITY]
    dev = ctx.upload_points(cases)
    want = ", ".join(f"[{x}, {y}, 1]" for x, y in cases)
    assert dev.text() == want
    assert dev.text(3, 4) == ", ".join(f"[{x}, {y}, 1]" for x, y in cases[3:7])
    assert dev.text(0, 0) == ""
    big = ctx.fixed_base(seed=0x5EEE, n=5000)
    assert big.text() == ", ".join(f"[{x}, {y}, 1]" for x, y in big.tolist())


def test_scalar_vectors_on_device(ctx, known_points):
    """vmsm_scalars_fold / _dot / _text_ptr and vmsm_msm_dev_ext: the witness / linear-form side of a folding round,
    against Python integers and the oracle."""
    from verifiable_mpc_b200 import VmsmError, _lib

    dlogs, pts = known_points
    L = E.L
    rng = random.Random(99)
    edge = [0, 1, L - 1, L >> 1, (L >> 1) + 1, 2**252, 2**252 - 1, 10**9, 10**9 - 1]
    for n in (2, 18, 1000, 70000):
        vals = (edge + [rng.randrange(L) for _ in range(n)])[:n]
        other = [rng.randrange(L) for _ in range(n)]
        a, b = ctx.upload_scalars(vals), ctx.upload_scalars(other)
        assert a.tolist() == vals
        want = ", ".join(str(v - L if v > (L >> 1) else v) for v in vals)
        assert a.text_bytes().decode() == want
        assert a.text_bytes(signed=False).decode() == ", ".join(map(str, vals))
        assert a.text_bytes(1, 1).decode() == str(vals[1] - L if vals[1] > (L >> 1) else vals[1])
        assert ctx.scalars_dot(a, 0, b, 0, n) == sum(x * y for x, y in zip(vals, other)) % L
        half = n // 2
        assert ctx.scalars_dot(a, half, b, 0, half) == sum(x * y for x, y in zip(vals[half:], other[:half])) % L
        c = rng.randrange(L)
        a.fold(half, c, _lib.FOLD_WITNESS)
        b.fold(half, c, _lib.FOLD_FORM)
        assert a.tolist(0, half) == [(x + c * y) % L for x, y in zip(vals[:half], vals[half:2 * half])]
        assert b.tolist(0, half) == [(c * x + y) % L for x, y in zip(other[:half], other[half:2 * half])]
        # element-wise forms on two vectors: z = r + c0*x, L_tilde = c1*L, and c*dst + src
        now_a, now_b = a.tolist(), b.tolist()
        a.axpy(c, b, _lib.AXPY_ADD_SCALED)
        assert a.tolist() == [(x + c * y) % L for x, y in zip(now_a, now_b)]
        b.axpy(c, None, _lib.AXPY_SCALE, off=1, n=n - 1)
        assert b.tolist() == now_b[:1] + [c * y % L for y in now_b[1:]]
        now_a, now_b = a.tolist(), b.tolist()
        a.axpy(L - 1, b, _lib.AXPY_SCALE_ADD)
        assert a.tolist() == [((L - 1) * x + y) % L for x, y in zip(now_a, now_b)]
        a.free(), b.free()
    assert ctx.scalars_dot(ctx.upload_scalars([5]), 0, ctx.upload_scalars([7]), 0, 0) == 0

    # A_i = g_R^{z_L} * k^{s}: z read in place at an offset, the extra scalar from the host, two MSMs in flight,
    # then the fold overwrites z while nothing may still be reading it
    n = 64
    g = ctx.upload_points(pts[:n])
    k = ctx.upload_points([pts[n]])
    z_vals = [rng.randrange(L) for _ in range(n)]
    z = ctx.upload_scalars(z_vals)
    half = n // 2
    for rnd in range(4):
        s_a, s_b = rng.randrange(L), rng.randrange(L)
        ctx.msm_dev_ext(g, half, half, z, 0, k, 0, [s_a], slot=0)
        ctx.msm_dev_ext(g, 0, half, z, half, k, 0, [s_b], slot=1)
        c = rng.randrange(L)
        z.fold(half, c, _lib.FOLD_WITNESS)  # issued while both MSMs are in flight
        assert ctx.result(0) == E.msm_naive(z_vals[:half] + [s_a], pts[half:n] + [pts[n]])
        assert ctx.result(1) == E.msm_naive(z_vals[half:] + [s_b], pts[:half] + [pts[n]])
        z_vals = [(x + c * y) % L for x, y in zip(z_vals[:half], z_vals[half:])] + z_vals[half:]
        assert z.tolist() == z_vals
    with pytest.raises(VmsmError):
        z.fold(n, 1, _lib.FOLD_WITNESS)
    with pytest.raises(VmsmError):
        z.fold(1, 1, 7)
    with pytest.raises(VmsmError):
        z.axpy(1, z, _lib.AXPY_ADD_SCALED, off=0, soff=1, n=8)  # overlapping ranges
    with pytest.raises(VmsmError):
        z.axpy(1, z, 9)
    with pytest.raises(VmsmError):
        ctx.scalars_dot(z, 1, z, 0, n)
    with pytest.raises(VmsmError):
        ctx.msm_dev_ext(g, 1, n, z, 0, k, 0, [1], slot=0)


def test_lincomb_async(ctx, known_points):
    """vmsm_lincomb_async: several small combinations in flight (more than the staging ring holds), results fetched
    out of order; an invalid input point surfaces as VMSM_ERR_POINT when its slot is fetched."""
    from verifiable_mpc_b200 import VmsmError

    dlogs, pts = known_points
    rng = random.Random(4)
    jobs = []
    for j in range(20):
        m = 1 + j % 5
        ps = [pts[rng.randrange(len(pts))] for _ in range(m)]
        sc = [rng.randrange(E.L) for _ in range(m)]
        ctx.lincomb_async(ps, sc, slot=j)
        jobs.append((ps, sc))
    for j in reversed(range(20)):
        ps, sc = jobs[j]
        assert ctx.result(j) == E.msm_naive(sc, ps), j
    ctx.lincomb_async([], [], slot=3)
    assert ctx.result(3) == E.IDENTITY
    ctx.lincomb_async([pts[0], (5, 7)], [1, 1], slot=4)  # (5, 7) is not on the curve
    with pytest.raises(VmsmError):
        ctx.result(4)
    ctx.lincomb_async([pts[0]], [2], slot=4)  # the slot is usable again
    assert ctx.result(4) == E.scalar_mul(pts[0], 2)


def test_abi_error_paths(ctx, known_points):
    """Status codes instead of crashes: bad handles, ranges, slots, options, mixed curves, Ed25519-only calls."""
    import ctypes

    from verifiable_mpc_b200 import VmsmError, _lib

    lib, h = ctx.lib, ctx.h
    _, pts = known_points
    dev = ctx.upload_points(pts[:8])
    out = ctypes.create_string_buffer(128)
    sc = b"\x01" + bytes(31)
    assert lib.vmsm_msm(h, 987654, 0, 1, sc, out) == _lib.ERR_INVALID
    assert lib.vmsm_msm(h, dev.handle, 8, 1, sc, out) == _lib.ERR_INVALID
    assert lib.vmsm_msm(h, dev.handle, 0, 1, None, out) == _lib.ERR_INVALID
    assert lib.vmsm_msm(12345, dev.handle, 0, 1, sc, out) == _lib.ERR_INVALID
    assert b"context" in lib.vmsm_last_error()
    assert lib.vmsm_points_download(h, dev.handle, 4, 5, out) == _lib.ERR_INVALID
    assert lib.vmsm_fold(h, dev.handle, 5, sc) == _lib.ERR_INVALID
    assert lib.vmsm_result_affine(h, 64, out) == _lib.ERR_INVALID
    assert lib.vmsm_ctx_set_option(h, 999, 1) == _lib.ERR_INVALID
    assert lib.vmsm_ctx_set_option(h, _lib.OPT_WINDOW_BITS, 40) == _lib.ERR_INVALID
    assert lib.vmsm_lincomb(h, 0, None, None, 65, out) == _lib.ERR_INVALID
    assert lib.vmsm_points_upload(h, 7, sc, 0, ctypes.byref(ctypes.c_uint64())) == _lib.ERR_UNSUPPORTED
    bn = ctx.fixed_base(seed=1, n=4, curve=1)
    assert lib.vmsm_fold(h, bn.handle, 2, sc) == _lib.ERR_UNSUPPORTED
    with pytest.raises(VmsmError):
        ctx.msm_ext(dev, 0, 2, bn, 0, 1, [1, 2, 3])  # mixed curves
    with pytest.raises(VmsmError):
        ctx.msm_dev_shard(dev, ctx.upload_scalars([1] * 8), slot=1, seq=1)  # no mailbox on this context
    # the context is still healthy afterwards
    assert ctx.msm(dev, [1] * 8) == E.msm_naive([1] * 8, pts[:8])
    dev.free()
    assert lib.vmsm_points_free(h, dev.handle) == _lib.ERR_INVALID


@pytest.mark.parametrize("logn", [8, 12, 14])
def test_back_to_back_msms_in_flight(ctx, logn):
    """Stress of the stream pipeline (sort stream / main stream / four tail streams): 40 different MSMs issued
    without any synchronisation in between, every result checked afterwards."""
    n = 1 << logn
    dev = ctx.fixed_base(seed=0x5EEE, n=n)
    dl = [prng.scalar(0x5EEE, i) for i in range(n)]
    nsets = 8
    scs, want = [], []
    for k in range(nsets):
        sc = [prng.scalar(0x3000 + k, i) for i in range(n)]
        scs.append(ctx.upload_scalars(sc))
        want.append(E.msm_known_dlog(sc, dl))
    for rep in range(3):
        for j in range(40):
            ctx.msm_dev(dev, scs[j % nsets], slot=j)
        for j in range(40):
            assert ctx.result(j) == want[j % nsets], (rep, j)


@pytest.mark.parametrize("sort_blocks", [0, 1, 7, 148, 100000])
def test_thin_counting_sort_grid(ctx, sort_blocks):
    """VMSM_OPT_SORT_BLOCKS: the per-scalar sort kernels of an MSM issued while the previous one is still accumulating
    run grid-stride with this many blocks; every setting must give the same results (2^14 terms = 64 full blocks)."""
    from verifiable_mpc_b200 import _lib

    n = 1 << 14
    dev = ctx.fixed_base(seed=0x5EEE, n=n)
    dl = [prng.scalar(0x5EEE, i) for i in range(n)]
    scs, want = [], []
    for k in range(3):
        sc = [prng.scalar(0x4000 + k, i) for i in range(n - k)]  # ragged lengths: the stride loop's tail
        scs.append(ctx.upload_scalars(sc))
        want.append(E.msm_known_dlog(sc, dl[:n - k]))
    ctx.set_option(_lib.OPT_SORT_BLOCKS, sort_blocks)
    try:
        for j in range(12):
            ctx.msm_dev(dev, scs[j % 3], slot=j, n=n - j % 3)
        for j in range(12):
            assert ctx.result(j) == want[j % 3], j
    finally:
        ctx.set_option(_lib.OPT_SORT_BLOCKS, -1)


def test_pooled_allocations_are_recycled_safely(ctx, known_points):
    """Freed point / scalar vectors go to a per-context pool and are handed out again (cudaFree would synchronise the
    device): many upload / free cycles of mixed sizes with MSMs in between must keep giving correct results, and a
    freed handle must stay invalid."""
    from verifiable_mpc_b200 import VmsmError

    dlogs, pts = known_points
    rng = random.Random(12)
    for it in range(60):
        n = rng.choice([1, 3, 17, 64, 200, 300])
        dev = ctx.upload_points(pts[:n])
        sc = [rng.randrange(E.L) for _ in range(n)]
        ds = ctx.upload_scalars(sc)
        ctx.msm_dev(dev, ds, slot=it % 8)
        got = ctx.result(it % 8)
        assert got == E.msm_known_dlog(sc, dlogs[:n]), it
        assert dev.tolist() == pts[:n] and ds.tolist() == sc
        handle = dev.handle
        dev.free()
        ds.free()
        if it == 0:
            stale = type(dev)(ctx, handle, n, 0)
            with pytest.raises(VmsmError):
                stale.tolist()
            stale.handle = 0


# ------------------------------------------------------------------------------------------------ precomputed bases
def test_msm_precomputed_small_matches_oracle(ctx, known_points):
    """vmsm_points_precompute at sizes the oracle's naive algorithm handles: every table window, every number of
    bucket sets, sub-ranges, extra terms (table built lazily), edge scalars, repeated / identity / negated bases."""
    from verifiable_mpc_b200 import _lib

    dl, pts = known_points
    ctx.set_option(_lib.OPT_PRE_MIN_TERMS, 0)
    try:
        n = 300
        scs = [prng.scalar(0x5EED, i) for i in range(n)]
        scs[:6] = [0, 1, E.L - 1, 2**252, 2**15, 2**16 - 1]
        want = E.msm_known_dlog(scs, dl[:n])
        for cb in (8, 11, 13, 16):
            dev = ctx.upload_points(pts[:n]).precompute(cb)
            W = (253 + cb) // cb
            for mode, seg_len in ((1, 0), (0, 0), (2, 3), (2, 64)):  # by geometry / per bucket / segments
                ctx.set_option(_lib.OPT_SEG_MODE, mode)
                ctx.set_option(_lib.OPT_SEG_LEN, seg_len)
                for sets in sorted({0, 1, 2, 3, W - 1, W}):
                    ctx.set_option(_lib.OPT_PRE_SETS, sets)
                    assert ctx.msm(dev, scs) == want, (cb, sets, mode, seg_len)
                    assert ctx.msm(dev, scs[:17]) == E.msm_naive(scs[:17], pts[:17]), (cb, sets, mode, seg_len)
            ctx.set_option(_lib.OPT_PRE_SETS, 0)
            ctx.set_option(_lib.OPT_SEG_MODE, 1)
            ctx.set_option(_lib.OPT_SEG_LEN, 0)
            # sub-range + extra terms: the one-element vector gets its table on first use
            extra = ctx.upload_points([pts[200], E.affine_neg(pts[201])])
            sc2 = scs[10:30] + [prng.scalar(9, 0), prng.scalar(9, 1)]
            assert ctx.msm_ext(dev, 7, 20, extra, 0, 2, sc2) == E.msm_known_dlog(sc2, dl[7:27] + [dl[200], E.L - dl[201]])
            dsc = ctx.upload_scalars(scs)
            ctx.msm_dev_ext(dev, 0, n, dsc, 0, extra, 1, [5], slot=9)
            assert ctx.result(9) == E.msm_known_dlog(scs + [5], dl[:n] + [E.L - dl[201]])
            # a fold drops the table and the folded vector still multiplies correctly
            c = prng.scalar(0xF01D, 1)
            dev.fold(c)
            assert not dev.precomputed
            fdl = [(c * dl[j] + dl[n // 2 + j]) % E.L for j in range(n // 2)]
            assert ctx.msm(dev, scs[:n // 2]) == E.msm_known_dlog(scs[:n // 2], fdl)
            dev.free(), extra.free(), dsc.free()
        pp = [pts[0]] * 5 + [E.IDENTITY] * 3 + [E.affine_neg(pts[0])] * 2
        sc = [5, 7, 11, 13, 17, 3, 4, 5, 9, 1]
        dev = ctx.upload_points(pp).precompute(8)
        assert ctx.msm(dev, sc) == E.msm_naive(sc, pp)
        assert ctx.msm(dev, [1, 1, 1, 1, 1, 0, 0, 0, 3, 2]) == E.IDENTITY
        dev.free()
    finally:
        ctx.set_option(_lib.OPT_PRE_MIN_TERMS, 256)
        ctx.set_option(_lib.OPT_PRE_SETS, 0)
        ctx.set_option(_lib.OPT_SEG_MODE, 1)
        ctx.set_option(_lib.OPT_SEG_LEN, 0)


@pytest.mark.parametrize("logn", [10, 12, 16, 18, 20])
def test_msm_precomputed_large_known_dlog(ctx, logn):
    """BASELINE sizes over precomputed bases: known-dlog identity, equality with the plain windowed path on the same
    inputs, skewed (boolean) scalars through the shared-bucket overflow path, several MSMs in flight."""
    from verifiable_mpc_b200 import _lib

    n = 1 << logn
    dev = ctx.fixed_base(seed=0x5EEE, n=n)
    sc = ctx.synth_scalars(0x5EED, n)
    ctx.msm_dev(dev, sc, slot=1)
    plain = ctx.result(1)
    assert plain == E.scalar_mul(E.B, _dlog_sum(0x5EED, 0x5EEE, n))
    dev.precompute()
    try:
        for mode in (1, 0, 2):  # accumulate kernel by geometry / one thread per bucket / equal segments
            ctx.set_option(_lib.OPT_SEG_MODE, mode)
            for sets in ((0, 1, 4) if logn <= 16 else (0, 8)):
                ctx.set_option(_lib.OPT_PRE_SETS, sets)
                for j in range(3):
                    ctx.msm_dev(dev, sc, slot=2 + j)
                assert [ctx.result(2 + j) for j in range(3)] == [plain] * 3, (mode, sets)
        ctx.set_option(_lib.OPT_PRE_SETS, 0)
        ctx.set_option(_lib.OPT_SEG_MODE, 1)
        m = min(n, 1 << 14)
        rnd = random.Random(logn)
        bits = [rnd.randrange(2) for _ in range(m)]
        dl = [prng.scalar(0x5EEE, i) for i in range(m)]
        assert ctx.msm(dev, bits) == E.msm_known_dlog(bits, dl)
        # sub-range starting in the middle of the vector
        off = n // 2 + 3
        sub = [prng.scalar(0x77, i) for i in range(min(1000, n - off))]
        assert ctx.msm(dev, sub, off=off) == E.msm_known_dlog(sub, [prng.scalar(0x5EEE, off + i) for i in range(len(sub))])
        # the plain path through the segmented kernel as well (boolean scalars: a bucket spanning thousands of segments)
        ctx.set_option(_lib.OPT_PRE_MIN_TERMS, 1 << 30)
        ctx.set_option(_lib.OPT_SEG_MODE, 2)
        ctx.msm_dev(dev, sc, slot=7)
        assert ctx.result(7) == plain
        assert ctx.msm(dev, bits) == E.msm_known_dlog(bits, dl)
    finally:
        ctx.set_option(_lib.OPT_PRE_SETS, 0)
        ctx.set_option(_lib.OPT_SEG_MODE, 1)
        ctx.set_option(_lib.OPT_PRE_MIN_TERMS, 256)
        dev.free()
        sc.free()


def test_fused_cross_term_commitment_and_host_normalisation(ctx, known_points):
    """vmsm_msm_dev_ext_dot: the scalar of the extra term is an inner product computed on the device (the cross terms
    of compressed_pivot.py:41-42), checked against the oracle; and VMSM_OPT_HOST_NORMALIZE on / off give the same
    canonical point (the CPU or a GPU thread inverts Z)."""
    from verifiable_mpc_b200 import _lib

    dl, pts = known_points
    n = 64
    dev = ctx.upload_points(pts[:2 * n])
    kd = ctx.upload_points([pts[250]])
    z = [prng.scalar(0xA1, i) for i in range(2 * n)]
    Lc = [prng.scalar(0xA2, i) for i in range(2 * n)]
    z[3], Lc[5], z[n + 1] = 0, E.L - 1, 1
    zd, Ld = ctx.upload_scalars(z), ctx.upload_scalars(Lc)
    s_a = sum(a * b for a, b in zip(Lc[n:], z[:n])) % E.L
    s_b = sum(a * b for a, b in zip(Lc[:n], z[n:])) % E.L
    want_a = E.msm_known_dlog(z[:n] + [s_a], dl[n:2 * n] + [dl[250]])
    want_b = E.msm_known_dlog(z[n:] + [s_b], dl[:n] + [dl[250]])
    try:
        for host_norm in (1, 0):
            ctx.set_option(_lib.OPT_HOST_NORMALIZE, host_norm)
            for rep in range(3):  # both staging parities, back to back
                ctx.msm_dev_ext_dot(dev, n, n, zd, 0, kd, 0, Ld, n, zd, 0, n, slot=0)
                ctx.msm_dev_ext_dot(dev, 0, n, zd, n, kd, 0, Ld, 0, zd, n, n, slot=1)
                assert ctx.result(0) == want_a and ctx.result(1) == want_b, (host_norm, rep)
            assert ctx.msm(dev, z) == E.msm_known_dlog(z, dl[:2 * n])
            assert ctx.lincomb(pts[:3], [5, E.L - 1, 0]) == E.msm_naive([5, E.L - 1, 0], pts[:3])
            assert ctx.msm(dev, [0] * 8) == E.IDENTITY
    finally:
        ctx.set_option(_lib.OPT_HOST_NORMALIZE, 1)
    with pytest.raises(Exception):
        ctx.msm_dev_ext_dot(dev, 0, n, zd, n, kd, 0, Ld, 0, zd, n + 1, n, slot=1)  # inner product range out of bounds
    for d in (dev, kd, zd, Ld):
        d.free()


@pytest.mark.parametrize("n", [1, 13, 1000, (1 << 12) - 5, 1 << 14])
def test_block_privatised_sort(ctx, n):
    """VMSM_OPT_BLOCK_SORT: digits recoded once into 16-bit codes, per-block shared-memory counters (vmsm_bsort_*).  With
    the size threshold lowered every geometry goes through it: plain windows (own bucket set each), forced windows up to
    c = 16 (128 KB of counters per block), tables with shared bucket sets, skewed scalars, ragged lengths, and MSMs issued
    back to back (256-thread blocks under the previous accumulate kernel, 1024-thread blocks alone).  Same results as
    the atomic two-pass sort and as the known-dlog identity."""
    from verifiable_mpc_b200 import _lib

    dev = ctx.fixed_base(seed=0x5EEE, n=n)
    dl = [prng.scalar(0x5EEE, i) for i in range(n)]
    rnd = random.Random(n)
    cases = [[prng.scalar(0xB50, i) for i in range(n)], [rnd.randrange(2) for _ in range(n)],
             [rnd.randrange(1, 4) << 240 for _ in range(n)]]
    cases[0][:3] = [0, E.L - 1, 1 << 15][:min(3, n)]
    want = [E.msm_known_dlog(sc, dl) for sc in cases]
    ups = [ctx.upload_scalars(sc) for sc in cases]
    try:
        ctx.set_option(_lib.OPT_BLOCK_SORT_MIN, 1)
        for bs in (1, 0):
            ctx.set_option(_lib.OPT_BLOCK_SORT, bs)
            for c in (0, 5, 12, 16):
                ctx.set_option(_lib.OPT_WINDOW_BITS, c)
                for j in range(9):  # back to back: the later ones sort underneath their predecessors
                    ctx.msm_dev(dev, ups[j % 3], slot=j)
                for j in range(9):
                    assert ctx.result(j) == want[j % 3], (bs, c, j)
        ctx.set_option(_lib.OPT_WINDOW_BITS, 0)
        ctx.set_option(_lib.OPT_BLOCK_SORT, 1)
        ctx.set_option(_lib.OPT_PRE_MIN_TERMS, 1)
        dev.precompute(16)
        for sets in (0, 1, 8):
            ctx.set_option(_lib.OPT_PRE_SETS, sets)
            for j in range(6):
                ctx.msm_dev(dev, ups[j % 3], slot=j)
            for j in range(6):
                assert ctx.result(j) == want[j % 3], (sets, j)
    finally:
        ctx.set_option(_lib.OPT_WINDOW_BITS, 0)
        ctx.set_option(_lib.OPT_PRE_SETS, 0)
        ctx.set_option(_lib.OPT_PRE_MIN_TERMS, 256)
        ctx.set_option(_lib.OPT_BLOCK_SORT_MIN, 1 << 15)
        ctx.set_option(_lib.OPT_BLOCK_SORT, 0)
        dev.free()
        for u in ups:
            u.free()
