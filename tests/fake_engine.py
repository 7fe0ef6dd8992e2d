"""Test-only stand-in for verifiable_mpc_b200.engine.Context that computes with the CPU oracle.

It lets the `-m "not gpu"` suite exercise the HOST logic of the prover twins (transcript layout, form algebra,
slicing / folding bookkeeping, PRNG draw order) against the golden fixtures in a container without a GPU.  It is
never importable from the product: the product's Context is libvmsm.so or nothing.
"""
from oracle import bn256 as BN
from oracle import ed25519 as E


class FakePoints:
    def __init__(self, ctx, pts):
        self.ctx, self.pts, self.handle, self.curve = ctx, list(pts), 1, 0

    @property
    def n(self):
        return len(self.pts)

    @n.setter
    def n(self, v):
        del self.pts[v:]

    def __len__(self):
        return len(self.pts)

    def download(self, off=0, n=None):
        n = len(self.pts) - off if n is None else n
        return b"".join(E.point_to_bytes(p) for p in self.pts[off:off + n])

    def tolist(self, off=0, n=None):
        n = len(self.pts) - off if n is None else n
        return self.pts[off:off + n]

    precomputed = False

    def precompute(self, window_bits=0):
        self.precomputed = True  # values are representation independent: the fake has nothing to build
        return self

    def fold(self, c):
        self.pts = E.fold(self.pts[: 2 * (len(self.pts) // 2)], int(c) % E.L)
        self.precomputed = False
        return self


def _unpack_scalars(raw):
    raw = bytes(raw)
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


class FakeScalars:
    """Device-resident scalar vector modulo the Ed25519 group order, on Python ints."""

    def __init__(self, ctx, vals, order=E.L):
        self.ctx, self.vals, self.handle = ctx, [int(v) % order for v in vals], 1

    @property
    def n(self):
        return len(self.vals)

    def tolist(self, off=0, n=None):
        n = len(self.vals) - off if n is None else n
        return self.vals[off:off + n]

    def download(self, off=0, n=None):
        return b"".join(v.to_bytes(32, "little") for v in self.tolist(off, n))

    def fold(self, half, c, mode):
        c = int(c) % E.L
        lo, hi = self.vals[:half], self.vals[half:2 * half]
        assert len(hi) == half
        new = [(a + c * b) % E.L for a, b in zip(lo, hi)] if mode == 0 else [(c * a + b) % E.L for a, b in zip(lo, hi)]
        self.vals[:half] = new

    def axpy(self, c, src=None, mode=0, off=0, soff=0, n=None):
        c = int(c) % E.L
        n = len(self.vals) - off if n is None else n
        d = self.vals[off:off + n]
        s = src.vals[soff:soff + n] if src is not None else [0] * n
        assert len(d) == n and len(s) == n
        new = [(a + c * b) % E.L for a, b in zip(d, s)] if mode == 0 else \
            [(c * a + b) % E.L for a, b in zip(d, s)] if mode == 1 else [c * a % E.L for a in d]
        self.vals[off:off + n] = new

    def text_bytes(self, off=0, n=None, signed=True):
        vals = self.tolist(off, n)
        return ", ".join(str(v - E.L if signed and v > (E.L >> 1) else v) for v in vals).encode()

    def free(self):
        self.handle = 0


class FakeBNPoints:
    def __init__(self, ctx, pts, curve):
        self.ctx, self.pts, self.handle, self.curve = ctx, list(pts), 1, curve

    @property
    def n(self):
        return len(self.pts)

    def tolist(self, off=0, n=None):
        n = len(self.pts) - off if n is None else n
        return self.pts[off:off + n]

    def free(self):
        self.handle = 0


class FakeContext:
    calls = 0

    def upload_points(self, pts, curve=0):
        if curve:
            F = BN.FP2 if curve == 2 else BN.FP
            for p in pts:
                assert BN.on_curve(F, p)
            return FakeBNPoints(self, pts, curve)
        if isinstance(pts, (bytes, bytearray)):
            pts = [E.point_from_bytes(pts[i:i + 64]) for i in range(0, len(pts), 64)]
        for p in pts:
            assert E.on_curve(p)
        return FakePoints(self, pts)

    def fixed_base(self, scalars=None, seed=0, n=None, curve=0):
        sc = _unpack_scalars(scalars) if isinstance(scalars, (bytes, bytearray)) else [int(s) for s in scalars]
        if curve:
            F, G = (BN.FP2, BN.G2) if curve == 2 else (BN.FP, BN.G1)
            return FakeBNPoints(self, [BN.scalar_mul(F, G, s % BN.N) for s in sc], curve)
        return FakePoints(self, [E.scalar_mul(E.B, s) for s in sc])

    def msm_ext(self, points, off, n, extra, extra_off, n_extra, scalars):
        FakeContext.calls += 1
        sc = _unpack_scalars(scalars) if isinstance(scalars, (bytes, bytearray)) else [int(s) % E.L for s in scalars]
        assert len(sc) == n + n_extra
        bases = points.pts[off:off + n] + extra.pts[extra_off:extra_off + n_extra]
        assert len(bases) == n + n_extra
        return E.msm_naive(sc, bases)

    def msm(self, points, scalars, off=0, n=None):
        if getattr(points, "curve", 0):
            F = BN.FP2 if points.curve == 2 else BN.FP
            sc = _unpack_scalars(scalars)
            FakeContext.calls += 1
            return BN.msm_naive(F, sc, points.pts[off:off + len(sc)])
        sc = _unpack_scalars(scalars) if isinstance(scalars, (bytes, bytearray)) else [int(s) % E.L for s in scalars]
        return E.msm_naive(sc, points.pts[off:off + len(sc)])

    def upload_scalars(self, scalars, order=E.L):
        assert order in (E.L, BN.N)
        return FakeScalars(self, _unpack_scalars(scalars) if isinstance(scalars, (bytes, bytearray)) else scalars, order)

    def scalars_dot(self, a, aoff, b, boff, n):
        assert aoff + n <= a.n and boff + n <= b.n
        return sum(x * y for x, y in zip(a.vals[aoff:aoff + n], b.vals[boff:boff + n])) % E.L

    def msm_dev_ext(self, points, poff, n, scalars, soff, extra, extra_off, extra_scalars, slot=0):
        FakeContext.calls += 1
        assert poff + n <= points.n and soff + n <= scalars.n
        assert extra_off + len(extra_scalars) <= extra.n and len(extra_scalars) <= 64
        curve = getattr(points, "curve", 0)
        assert getattr(extra, "curve", 0) == curve
        sc = scalars.vals[soff:soff + n] + [int(s) % (BN.N if curve else E.L) for s in extra_scalars]
        bases = points.pts[poff:poff + n] + extra.pts[extra_off:extra_off + len(extra_scalars)]
        if not hasattr(self, "_slots"):
            self._slots = {}
        self._slots[slot] = BN.msm_naive(BN.FP2 if curve == 2 else BN.FP, sc, bases) if curve else E.msm_naive(sc, bases)

    def msm_dev_ext_dot(self, points, poff, n, scalars, soff, extra, extra_off, dot_a, dot_aoff, dot_b, dot_boff, dot_n,
                        slot=0):
        s = self.scalars_dot(dot_a, dot_aoff, dot_b, dot_boff, dot_n)
        self.msm_dev_ext(points, poff, n, scalars, soff, extra, extra_off, [s], slot=slot)

    def msm_dev(self, points, scalars, slot=0, poff=0, soff=0, n=None):
        FakeContext.calls += 1
        if n is None:
            n = min(points.n - poff, scalars.n - soff)
        if not hasattr(self, "_slots"):
            self._slots = {}
        curve = getattr(points, "curve", 0)
        sc, bases = scalars.vals[soff:soff + n], points.pts[poff:poff + n]
        self._slots[slot] = BN.msm_naive(BN.FP2 if curve == 2 else BN.FP, sc, bases) if curve else E.msm_naive(sc, bases)

    def msm_async(self, points, ptr, off, n, slot):
        import ctypes

        raw = ctypes.string_at(ptr, 32 * n) if n else b""
        if not hasattr(self, "_slots"):
            self._slots = {}
        self._slots[slot] = self.msm(points, raw, off=off)

    def result(self, slot=0, curve=0):
        return self._slots[slot]

    def concat(self, a, a_off, a_n, b=None, b_off=0, b_n=0):
        pts = a.pts[a_off:a_off + a_n] + (b.pts[b_off:b_off + b_n] if b is not None else [])
        return FakePoints(self, pts)

    def lincomb_async(self, pts, scalars, slot, curve=0):
        assert curve == 0
        if not hasattr(self, "_slots"):
            self._slots = {}
        self._slots[slot] = self.lincomb(pts, scalars)

    def lincomb(self, pts, scalars, curve=0):
        FakeContext.calls += 1
        if curve:
            F = BN.FP2 if curve == 2 else BN.FP
            return BN.msm_naive(F, [int(s) % BN.N for s in scalars], list(pts))
        return E.msm_naive([int(s) % E.L for s in scalars], list(pts)) if pts else E.IDENTITY
