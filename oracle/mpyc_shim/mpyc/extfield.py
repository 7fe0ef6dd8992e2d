"""GF(p^2) = GF(p)[x]/(x^2+1) look-alike of MPyC's extension fields, just enough for verifiable_mpc/ac20/pairing.py to
IMPORT (it evaluates xi ** ((p-1)//6) etc. at module load, pairing.py:55-78) and for the BN256_twist group type's
`.field`.  Elements are built from little-endian coefficient lists (`GFp_2([0, 1, 0])` is x, pairing.py:55) or ints and
expose `.value.value[k]` (pairing.py:73-74).  TEST INFRASTRUCTURE ONLY."""
import functools

from oracle import bn256 as _bn


class _Poly:
    __slots__ = ("value",)

    def __init__(self, coeffs):
        self.value = coeffs  # [c0, c1] ints mod p


class ExtensionFieldElement:
    __slots__ = ("value",)
    modulus = None
    order = None
    characteristic = _bn.P
    ext_deg = 2
    is_signed = False

    def __init__(self, value=0):
        P = _bn.P
        if isinstance(value, ExtensionFieldElement):
            c = list(value.value.value)
        elif isinstance(value, (list, tuple)):
            c = [int(getattr(v, "value", v)) % P for v in value]
            assert not any(c[2:]), "degree < 2 expected"
            c = (c + [0, 0])[:2]
        else:
            c = [int(getattr(value, "value", value)) % P, 0]
        self.value = _Poly(c)

    def _pair(self):
        return (self.value.value[0], self.value.value[1])

    @classmethod
    def _from_pair(cls, t):
        return cls([t[0], t[1]])

    def _coerce(self, other):
        if isinstance(other, ExtensionFieldElement):
            return other._pair()
        if isinstance(other, int):
            return (other % _bn.P, 0)
        v = getattr(other, "value", None)
        if isinstance(v, int):
            return (v % _bn.P, 0)
        return None

    def __add__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else self._from_pair(_bn.f2_add(self._pair(), o))

    __radd__ = __add__

    def __sub__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else self._from_pair(_bn.f2_sub(self._pair(), o))

    def __rsub__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else self._from_pair(_bn.f2_sub(o, self._pair()))

    def __neg__(self):
        return self._from_pair(_bn.f2_sub((0, 0), self._pair()))

    def __mul__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else self._from_pair(_bn.f2_mul(self._pair(), o))

    __rmul__ = __mul__

    def reciprocal(self):
        return self._from_pair(_bn.f2_inv(self._pair()))

    def __truediv__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else self._from_pair(_bn.f2_mul(self._pair(), _bn.f2_inv(o)))

    def __pow__(self, e):
        e = int(e)
        base, acc = self._pair(), (1, 0)
        if e < 0:
            base, e = _bn.f2_inv(base), -e
        while e:
            if e & 1:
                acc = _bn.f2_mul(acc, base)
            base = _bn.f2_mul(base, base)
            e >>= 1
        return self._from_pair(acc)

    def __eq__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else self._pair() == o

    def __hash__(self):
        return hash(self._pair())

    def __int__(self):
        # integer encoding of the coefficient polynomial (c0 + c1 * p), zero iff the element is zero
        return self.value.value[0] + self.value.value[1] * _bn.P

    def __bool__(self):
        return any(self.value.value)

    def __repr__(self):
        return repr(list(self.value.value))


@functools.lru_cache(maxsize=None)
def bn256_fp2():
    cls = type("GF(p^2)", (ExtensionFieldElement,), {"__slots__": ()})
    cls.order = _bn.P ** 2
    cls.modulus = "x^2+1"
    return cls


def ext_field(modulus):
    return bn256_fp2()
