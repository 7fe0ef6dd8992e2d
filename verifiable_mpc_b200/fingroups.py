"""MPyC-style group-element types whose arithmetic runs on the B200 (north_star: "keep ... MPyC group-element types").

``EllipticCurve('Ed25519', 'projective')`` returns a class with the surface the reference uses on
``mpyc.fingroups.EllipticCurve`` classes: ``.order .generator .identity .field``, class-level
``is_additive`` / ``is_multiplicative`` flags (set by the caller: demos/demo_zkp_ac20.py:46-48), instances with
``a * b``, ``a ** n`` (also negative / huge n), ``n * a``, ``a + b``, ``a @ b``, ``~a``, ``==``, ``.normalize()``,
``.x .y .z``, ``type(a).operation``.

Every group operation is a call into libvmsm.so (``Context.lincomb``); points are held on the host as canonical
affine integers, so ``repr()`` is representation independent: ``[x, y, 1]`` with unsigned decimal coordinates.
Single operations cost a kernel launch each -- they are meant for the O(log N) glue of a proof
(``Q' = A * Q**c * B**(c**2)``); anything O(N) goes through ``DevicePointList`` (device-resident generator vectors).
"""
import functools

from . import _lib
from .engine import BN_N, BN_P, ED_L, ED_P, default_context, unpack_points
from .finfields import GF

BX = 15112221349535400772501151409588531511454012693041857206046113283949847762202
BY = 46316835694926478169428394003475163141307993866256225615783033603165251855960


class FiniteGroupElement:
    __slots__ = ()
    order = None
    is_additive = False
    is_multiplicative = False
    is_abelian = True
    identity = None
    generator = None

    def __matmul__(self, other):
        if not isinstance(other, type(self)):
            return NotImplemented
        return type(self).operation(self, other)

    def __invert__(self):
        return type(self).inversion(self)

    def __xor__(self, n):
        return type(self).repeat(self, int(n))

    def __mul__(self, other):
        cls = type(self)
        if cls.is_multiplicative and isinstance(other, cls):
            return cls.operation(self, other)
        if cls.is_additive and isinstance(other, int):
            return cls.repeat(self, other)
        return NotImplemented

    def __rmul__(self, other):
        cls = type(self)
        if cls.is_additive and isinstance(other, int):
            return cls.repeat(self, other)
        if cls.is_multiplicative and isinstance(other, cls):
            return cls.operation(other, self)
        return NotImplemented

    def __truediv__(self, other):
        cls = type(self)
        if cls.is_multiplicative and isinstance(other, cls):
            return cls.operation(self, cls.inversion(other))
        return NotImplemented

    def __pow__(self, n):
        cls = type(self)
        if not cls.is_multiplicative:
            raise TypeError("group not multiplicative")
        return cls.repeat(self, int(n))

    def __add__(self, other):
        cls = type(self)
        if cls.is_additive and isinstance(other, cls):
            return cls.operation(self, other)
        return NotImplemented

    def __sub__(self, other):
        cls = type(self)
        if cls.is_additive and isinstance(other, cls):
            return cls.operation(self, cls.inversion(other))
        return NotImplemented

    def __neg__(self):
        cls = type(self)
        if not cls.is_additive:
            raise TypeError("group not additive")
        return cls.inversion(self)


class EllipticCurvePoint(FiniteGroupElement):
    __slots__ = ()
    field = None


class Ed25519Point(EllipticCurvePoint):
    """Canonical affine (x, y) on the host; arithmetic on the device."""
    __slots__ = ("ax", "ay")
    order = ED_L
    curvename = "Ed25519"
    context = None  # engine.Context used by this class; defaults to the process-wide one

    def __init__(self, value=None, check=True):
        if value is None:
            value = (0, 1, 1)
        vals = [int(v) for v in value]
        if len(vals) == 2:
            vals.append(1)
        X, Y, Z = (v % ED_P for v in vals[:3])
        if Z != 1:
            zi = pow(Z, -1, ED_P)
            X, Y = X * zi % ED_P, Y * zi % ED_P
        self.ax, self.ay = X, Y
        if check:
            d = (-121665 * pow(121666, -1, ED_P)) % ED_P
            assert (-X * X + Y * Y - 1 - d * X * X * Y * Y) % ED_P == 0, "point not on curve"

    @classmethod
    def _ctx(cls):
        return cls.context or default_context()

    @classmethod
    def _make(cls, xy):
        obj = cls.__new__(cls)
        obj.ax, obj.ay = xy
        return obj

    # -- MPyC surface
    @property
    def value(self):
        F = type(self).field
        return [F(self.ax), F(self.ay), F(1)]

    @property
    def x(self):
        return type(self).field(self.ax)

    @property
    def y(self):
        return type(self).field(self.ay)

    @property
    def z(self):
        return type(self).field(1)

    def affine(self):
        return (self.ax, self.ay)

    def normalize(self):
        return self

    def __repr__(self):
        return f"[{self.ax}, {self.ay}, 1]"

    def __eq__(self, other):
        if not isinstance(other, type(self)):
            return NotImplemented
        return self.ax == other.ax and self.ay == other.ay

    def __hash__(self):
        return hash((self.ax, self.ay))

    @classmethod
    def operation(cls, a, b):
        return cls._make(cls._ctx().lincomb([a.affine(), b.affine()], [1, 1]))

    @classmethod
    def operation2(cls, a):
        return cls._make(cls._ctx().lincomb([a.affine()], [2]))

    @classmethod
    def inversion(cls, a):
        return cls._make(((-a.ax) % ED_P, a.ay))

    @classmethod
    def equality(cls, a, b):
        return a == b

    @classmethod
    def repeat(cls, a, n):
        # exponents arrive negative and unreduced (pivot.py:187, compressed_pivot.py:66); every element the
        # provers handle lies in the order-l subgroup generated by B, so reduction mod l is exact
        return cls._make(cls._ctx().lincomb([a.affine()], [int(n) % ED_L]))

    @classmethod
    def lincomb(cls, points, scalars):
        """prod_i points[i] ** scalars[i] in ONE device call (<= 64 terms)."""
        return cls._make(cls._ctx().lincomb([p.affine() for p in points], [int(s) % ED_L for s in scalars]))


@functools.lru_cache(maxsize=None)
def _ed25519_class(coordinates):
    F = GF(ED_P)
    cls = type(f"E({F.__name__}){coordinates}", (Ed25519Point,), {"__slots__": ()})
    cls.field = F
    cls.coordinates = coordinates
    cls.is_additive = True
    cls.is_multiplicative = False
    cls.identity = cls._make((0, 1))
    cls.generator = cls._make((BX, BY))
    return cls


# ---------------------------------------------------------------------------------------------- BN256 G1 / G2
class BN256Point(EllipticCurvePoint):
    """A point of BN256 (G1, over Fp) or of its sextic twist (G2, over Fp2 = Fp[i]/(i^2+1)): the groups the
    Pinocchio prover works in (demos/demo_zkp_pynocchio.py:27-29).  Host side: ``pt`` is None (identity), ``(x, y)``
    for G1 or ``((x_re, x_im), (y_re, y_im))`` for G2, canonical; every group operation is a device call."""
    __slots__ = ("pt",)
    order = BN_N
    curve_id = _lib.CURVE_BN256_G1
    context = None

    def __init__(self, value=None, check=True):
        if value is None:
            self.pt = None
            return
        raise NotImplementedError("construct BN256 points with from_affine() or from group.generator")

    @classmethod
    def _ctx(cls):
        return cls.context or default_context()

    @classmethod
    def _make(cls, pt):
        obj = cls.__new__(cls)
        obj.pt = pt
        return obj

    from_affine = _make

    def affine(self):
        return self.pt

    def normalize(self):
        return self

    def __repr__(self):
        if self.pt is None:
            return "[1, 1, 0]"
        x, y = self.pt
        return f"[{list(x) if isinstance(x, tuple) else x}, {list(y) if isinstance(y, tuple) else y}, 1]"

    def __eq__(self, other):
        if not isinstance(other, type(self)):
            return NotImplemented
        return self.pt == other.pt

    def __hash__(self):
        return hash(self.pt)

    @classmethod
    def operation(cls, a, b):
        return cls._make(cls._ctx().lincomb([a.pt, b.pt], [1, 1], curve=cls.curve_id))

    @classmethod
    def operation2(cls, a):
        return cls._make(cls._ctx().lincomb([a.pt], [2], curve=cls.curve_id))

    @classmethod
    def inversion(cls, a):
        if a.pt is None:
            return a
        x, y = a.pt
        ny = tuple((-v) % BN_P for v in y) if isinstance(y, tuple) else (-y) % BN_P
        return cls._make((x, ny))

    @classmethod
    def equality(cls, a, b):
        return a == b

    @classmethod
    def repeat(cls, a, n):
        return cls._make(cls._ctx().lincomb([a.pt], [int(n) % BN_N], curve=cls.curve_id))

    @classmethod
    def lincomb(cls, points, scalars):
        return cls._make(cls._ctx().lincomb([p.pt for p in points], [int(s) % BN_N for s in scalars], curve=cls.curve_id))


_BN_G1 = (1, BN_P - 2)
_BN_G2 = ((64746500191241794695844075326670126197795977525365406531717464316923369116492,
           21167961636542580255011770066570541300993051739349375019639421053990175267184),
          (17778617556404439934652658462602675281523610326338642107814333856843981424549,
           20666913350058776956210519119118544732556678129809273996262322366050359951122))


@functools.lru_cache(maxsize=None)
def _bn256_class(curvename, coordinates):
    twist = curvename == "BN256_twist"
    cls = type(f"E({curvename}){coordinates}", (BN256Point,), {"__slots__": ()})
    cls.curve_id = _lib.CURVE_BN256_G2 if twist else _lib.CURVE_BN256_G1
    cls.curvename = curvename
    cls.coordinates = coordinates
    cls.field = None if twist else GF(BN_P)
    cls.is_additive, cls.is_multiplicative = True, False
    cls.identity = cls._make(None)
    cls.generator = cls._make(_BN_G2 if twist else _BN_G1)
    return cls


def EllipticCurve(curvename="Ed25519", coordinates=None):
    if curvename == "Ed25519":
        return _ed25519_class(coordinates or "extended")
    if curvename in ("BN256", "BN256_twist"):
        return _bn256_class(curvename, coordinates or "jacobian")
    raise NotImplementedError(f"curve {curvename} is not part of the engine")


class DevicePointList:
    """A list-like view [off, off+n) of a device-resident point vector (generator lists g, g_hat).

    Behaves like the Python lists of group elements the reference passes around -- ``len``, indexing, slicing,
    iteration, ``repr`` (byte-identical to ``repr`` of the list of elements, which is what enters the Fiat-Shamir
    pre-image, compressed_pivot.py:51-59) -- while the points themselves stay in HBM for the MSM / fold kernels.
    """

    def __init__(self, group, dev, off=0, n=None):
        self.group, self.dev, self.off = group, dev, off
        self.n = dev.n - off if n is None else n

    @classmethod
    def from_points(cls, group, points, ctx=None):
        ctx = ctx or group._ctx()
        return cls(group, ctx.upload_points([p.affine() for p in points]))

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            start, stop, step = i.indices(self.n)
            if step != 1:
                raise ValueError("only contiguous slices")
            return DevicePointList(self.group, self.dev, self.off + start, max(0, stop - start))
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return self.group._make(self.dev.tolist(self.off + i, 1)[0])

    def precompute(self, window_bits=0):
        """Fixed generators: build the table of ``2^(c*w) * g_i`` once (``vmsm_points_precompute``); every later
        ``pivot.vector_commitment`` on this list -- the z commitment and the announcement A of each proof
        (circuit_sat_cb.py:103, compressed_pivot.py:110) -- then runs without doublings."""
        self.dev.precompute(window_bits)
        return self

    def affine_list(self):
        return unpack_points(self.dev.download(self.off, self.n))

    def __iter__(self):
        return (self.group._make(xy) for xy in self.affine_list())

    def tolist(self):
        return list(self)

    def __repr__(self):
        # byte-identical to repr(list of points).  The decimal text is produced on the device (vmsm_points_text);
        # formatting 2n 255-bit integers in Python used to dominate the prover's latency.
        if hasattr(self.dev, "text"):
            return "[" + self.dev.text(self.off, self.n) + "]"
        raw = self.dev.download(self.off, self.n)
        fb = int.from_bytes
        return "[" + ", ".join([f"[{fb(raw[i:i + 32], 'little')}, {fb(raw[i + 32:i + 64], 'little')}, 1]"
                                for i in range(0, len(raw), 64)]) + "]"

    def repr_bytes(self):
        """repr(self).encode() without ever building the str (fed straight into the incremental hash)."""
        if hasattr(self.dev, "text_bytes"):
            return b"[" + self.dev.text_bytes(self.off, self.n) + b"]"
        return repr(self).encode("utf-8")

    def prefetch_repr(self):
        """Produce the text of the current view now (device kernels + copy into the context's pinned buffer) and keep
        the view for the next ``feed_repr``.  The prover calls this while the commitments that precede the generators in
        the round's pre-image are still being computed; the view is dropped by anything that changes the vector."""
        if hasattr(self.dev, "text_view"):
            self._text = ((self.dev.handle, self.off, self.n), self.dev.text_view(self.off, self.n))

    def feed_repr(self, h):
        """h.update(repr(self).encode()) with the text hashed IN PLACE from the context's pinned buffer: no Python
        bytes object of the ~160 bytes per point is ever built (two 10 MB copies per round at N = 2^16 otherwise)."""
        if hasattr(self.dev, "text_view"):
            key, view = getattr(self, "_text", None) or (None, None)
            self._text = None  # single use: the pinned buffer belongs to the next text call
            if key != (self.dev.handle, self.off, self.n):
                view = self.dev.text_view(self.off, self.n)
            h.update(b"[")
            h.update(view)
            h.update(b"]")
        else:
            h.update(self.repr_bytes())

    def wire_bytes(self):
        """Canonical 64-byte encodings x || y of the whole view (binary transcript mode); a zero-copy view of the
        context's pinned buffer when the engine offers one."""
        if hasattr(self.dev, "wire_view"):
            return self.dev.wire_view(self.off, self.n)
        return self.dev.download(self.off, self.n)

    def __eq__(self, other):
        if isinstance(other, DevicePointList):
            return self.affine_list() == other.affine_list()
        if isinstance(other, list):
            return self.affine_list() == [p.affine() for p in other]
        return NotImplemented

    def clone(self, extra=None):
        """Private device copy, optionally with one more vector appended (``g + [h]``)."""
        ctx = self.dev.ctx
        if extra is None:
            return DevicePointList(self.group, ctx.concat(self.dev, self.off, self.n))
        return DevicePointList(self.group, ctx.concat(self.dev, self.off, self.n, extra.dev, extra.off, extra.n))
