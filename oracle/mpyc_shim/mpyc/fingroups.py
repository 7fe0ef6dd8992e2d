"""mpyc.fingroups look-alike on the pure-Python oracle (TEST INFRASTRUCTURE ONLY).

Surface used by the reference: EllipticCurve(name, coordinates) -> class with .order .generator
.identity .field and class-level flags is_additive / is_multiplicative; instances with .x .y .z, normalize(), ~a,
a * b, a ** n (multiplicative), n * a, a + b (additive), a @ b, == ; FiniteGroupElement.__matmul__; QuadraticResidues.

UNVERIFIED against real MPyC (source absent here): repeat() is a right-to-left binary double-and-add, Edwards
'projective' uses the EFD bbjlp formulas, repr() convention (see below).  Group VALUES are representation independent.
repr(point) := "[x, y, 1]" of the canonical affine point with unsigned decimal coordinates (deliberate: transcripts
must not depend on the projective representative, SURVEY.md hard part 1).
"""
import functools

from oracle import ed25519 as _ed

from .finfields import GF


class FiniteGroupElement:
    __slots__ = ("value",)
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

    # multiplicative notation
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

    # additive notation
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

    def __eq__(self, other):
        if not isinstance(other, type(self)):
            return NotImplemented
        return type(self).equality(self, other)

    def __hash__(self):
        return hash(repr(self))

    @classmethod
    def repeat(cls, a, n):
        """Right-to-left binary double-and-add (generic fallback)."""
        if n < 0:
            a = cls.inversion(a)
            n = -n
        if n == 0:
            return cls.identity
        d, c = a, cls.identity
        for i in range(n.bit_length() - 1):
            if (n >> i) & 1:
                c = cls.operation(c, d)
            d = cls.operation2(d)
        return cls.operation(c, d)

    @classmethod
    def operation2(cls, a):
        return cls.operation(a, a)


class EllipticCurvePoint(FiniteGroupElement):
    __slots__ = ()
    field = None
    is_additive = True

    @property
    def x(self):
        return self.value[0]

    @property
    def y(self):
        return self.value[1]

    @property
    def z(self):
        return self.value[2]


# ------------------------------------------------------------------------------------------------ Ed25519
class _Ed25519Base(EllipticCurvePoint):
    """value = [X, Y, Z] (or [X, Y] for affine) of field elements; arithmetic runs on raw ints via oracle.ed25519."""
    __slots__ = ()
    order = _ed.L
    coordinates = "projective"

    def __init__(self, value=None, check=True):
        F = type(self).field
        if value is None:
            value = (0, 1, 1)
        value = [v if isinstance(v, F) else F(int(v)) for v in value]
        if len(value) == 2:
            value.append(F(1))
        self.value = value
        if check:
            X, Y, Z = (int(v.value) for v in value)
            zi = pow(Z, -1, _ed.P)
            assert _ed.on_curve((X * zi % _ed.P, Y * zi % _ed.P)), "point not on curve"

    def _ints(self):
        return tuple(v.value for v in self.value)

    @classmethod
    def _from_ints(cls, t):
        obj = cls.__new__(cls)
        F = cls.field
        obj.value = [F(t[0]), F(t[1]), F(t[2])]
        return obj

    def normalize(self):
        x, y = _ed.normalize(self._ints())
        return type(self)._from_ints((x, y, 1))

    def affine(self):
        return _ed.normalize(self._ints())

    def __repr__(self):
        x, y = _ed.normalize(self._ints())
        return f"[{x}, {y}, 1]"

    @classmethod
    def operation(cls, a, b):
        return cls._from_ints(_ed.proj_add(a._ints(), b._ints()))

    @classmethod
    def operation2(cls, a):
        return cls._from_ints(_ed.proj_dbl(a._ints()))

    @classmethod
    def inversion(cls, a):
        return cls._from_ints(_ed.proj_neg(a._ints()))

    @classmethod
    def equality(cls, a, b):
        return _ed.proj_eq(a._ints(), b._ints())

    @classmethod
    def repeat(cls, a, n):
        return cls._from_ints(_ed.proj_repeat(a._ints(), int(n)))


@functools.lru_cache(maxsize=None)
def _ed25519_class(coordinates):
    F = GF(_ed.P)
    cls = type(f"E({F.__name__}){coordinates}", (_Ed25519Base,), {"__slots__": ()})
    cls.field = F
    cls.coordinates = coordinates
    cls.is_additive = True
    cls.is_multiplicative = False
    cls.identity = cls._from_ints((0, 1, 1))
    cls.generator = cls._from_ints((_ed.BX, _ed.BY, 1))
    cls.curvename = "Ed25519"
    return cls


def EllipticCurve(curvename="Ed25519", coordinates=None):
    if curvename == "Ed25519":
        return _ed25519_class(coordinates or "extended")
    if curvename in ("BN256", "BN256_twist"):
        from . import bn256_groups

        return bn256_groups.curve_class(curvename, coordinates or "jacobian")
    raise ValueError(f"curve {curvename} not in the test shim")


# ------------------------------------------------------------------------------------------------ QR groups
def _is_prime(n):
    if n < 2:
        return False
    for q in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % q == 0:
            return n == q
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


@functools.lru_cache(maxsize=None)
def _safe_prime(l):
    q = (1 << (l - 2)) | 1
    while True:
        if q % 3 != 0 and _is_prime(q) and _is_prime(2 * q + 1):
            return 2 * q + 1
        q += 2


class _QRBase(FiniteGroupElement):
    __slots__ = ()
    is_multiplicative = True
    modulus = None

    def __init__(self, value=1, check=True):
        self.value = int(value) % type(self).modulus

    def __int__(self):
        return self.value

    def __repr__(self):
        return repr(self.value)

    @classmethod
    def operation(cls, a, b):
        return cls(a.value * b.value)

    @classmethod
    def inversion(cls, a):
        return cls(pow(a.value, -1, cls.modulus))

    @classmethod
    def equality(cls, a, b):
        return a.value == b.value

    @classmethod
    def repeat(cls, a, n):
        return cls(pow(a.value, int(n), cls.modulus))


@functools.lru_cache(maxsize=None)
def QuadraticResidues(p=None, l=None):
    if p is None:
        p = _safe_prime(l or 2048)
    cls = type(f"QR({p})", (_QRBase,), {"__slots__": ()})
    cls.modulus = p
    cls.order = (p - 1) // 2
    cls.identity = cls(1)
    cls.generator = cls(4)
    cls.field = GF(p)
    return cls
