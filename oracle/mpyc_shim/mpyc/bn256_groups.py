"""BN256 / BN256_twist group types of the MPyC look-alike (TEST INFRASTRUCTURE ONLY), on oracle/bn256.py.
Additive notation by default, as MPyC's elliptic-curve groups; `int * pt`, `pt + pt`, `-pt`, `a @ b`, `==`,
`.normalize()`, `.x .y .z`, `type.field/.order/.generator/.identity`."""
import functools

from oracle import bn256 as _bn

from .finfields import GF
from .fingroups import EllipticCurvePoint
from . import extfield


class _BNBase(EllipticCurvePoint):
    __slots__ = ()
    order = _bn.N
    F = None        # oracle field view
    coordinates = "jacobian"

    def __init__(self, value=None, check=True):
        fld = type(self).field
        if value is None:
            self.value = [fld(1), fld(1), fld(0)]
            return
        value = [v if isinstance(v, fld) else fld(v) for v in value]
        if len(value) == 2:
            value.append(fld(1))
        self.value = value
        if check:
            assert _bn.on_curve(type(self).F, self.affine()), "point not on curve"

    def _raw(self, v):
        return (v.value.value[0], v.value.value[1]) if type(self).F.ext else v.value

    def _jac(self):
        return tuple(self._raw(v) for v in self.value)

    @classmethod
    def _from_jac(cls, j):
        obj = cls.__new__(cls)
        fld = cls.field
        obj.value = [fld(list(c)) if cls.F.ext else fld(c) for c in j]
        return obj

    @classmethod
    def from_affine(cls, pt):
        return cls._from_jac(_bn.to_jac(cls.F, pt))

    def affine(self):
        return _bn.normalize(type(self).F, self._jac())

    def normalize(self):
        return type(self).from_affine(self.affine())

    def __repr__(self):
        return repr(self.affine())

    @classmethod
    def operation(cls, a, b):
        return cls._from_jac(_bn.jac_add(cls.F, a._jac(), b._jac()))

    @classmethod
    def operation2(cls, a):
        return cls._from_jac(_bn.jac_dbl(cls.F, a._jac()))

    @classmethod
    def inversion(cls, a):
        X, Y, Z = a._jac()
        return cls._from_jac((X, cls.F.sub(cls.F.zero, Y), Z))

    @classmethod
    def equality(cls, a, b):
        return a.affine() == b.affine()

    @classmethod
    def repeat(cls, a, n):
        return cls._from_jac(_bn.jac_repeat(cls.F, a._jac(), int(n)))


@functools.lru_cache(maxsize=None)
def curve_class(curvename, coordinates):
    twist = curvename == "BN256_twist"
    cls = type(f"E({curvename}){coordinates}", (_BNBase,), {"__slots__": ()})
    cls.F = _bn.FP2 if twist else _bn.FP
    cls.field = extfield.bn256_fp2() if twist else GF(_bn.P)
    cls.coordinates = coordinates
    cls.curvename = curvename
    cls.is_additive, cls.is_multiplicative = True, False
    cls.identity = cls.from_affine(None)
    cls.generator = cls.from_affine(_bn.G2 if twist else _bn.G1)
    return cls
