"""mpyc.finfields look-alike: GF(modulus) prime fields (+ GF(p^2) as polynomial extension, see extfield below).

Behaviour relied upon by the reference: class attributes .order/.modulus/.is_signed, callable on
ints, + - * / ** with ints, int(e) (signed representative when is_signed), .value, == int, hash.
"""
import functools


class FiniteFieldElement:
    __slots__ = ("value",)
    modulus = None
    order = None
    is_signed = True


class PrimeFieldElement(FiniteFieldElement):
    __slots__ = ()

    def __init__(self, value=0):
        if isinstance(value, FiniteFieldElement):
            value = value.value
        self.value = int(value) % type(self).modulus

    # conversions
    def __int__(self):
        v = self.value
        if type(self).is_signed and v > type(self).modulus >> 1:
            return v - type(self).modulus
        return v

    __index__ = __int__

    def __repr__(self):
        return f"{int(self)}"

    def __hash__(self):
        return hash((type(self).__name__, self.value))

    def __bool__(self):
        return self.value != 0

    def _coerce(self, other):
        if isinstance(other, PrimeFieldElement):
            if type(other) is not type(self):
                return None
            return other.value
        if isinstance(other, int):
            return other
        return None

    def __eq__(self, other):
        o = self._coerce(other)
        if o is None:
            return NotImplemented
        return self.value == o % type(self).modulus

    def __add__(self, other):
        o = self._coerce(other)
        if o is None:
            return NotImplemented
        return type(self)(self.value + o)

    __radd__ = __add__

    def __sub__(self, other):
        o = self._coerce(other)
        if o is None:
            return NotImplemented
        return type(self)(self.value - o)

    def __rsub__(self, other):
        o = self._coerce(other)
        if o is None:
            return NotImplemented
        return type(self)(o - self.value)

    def __mul__(self, other):
        o = self._coerce(other)
        if o is None:
            return NotImplemented
        return type(self)(self.value * o)

    __rmul__ = __mul__

    def __neg__(self):
        return type(self)(-self.value)

    def __pos__(self):
        return self

    def reciprocal(self):
        return type(self)(pow(self.value, -1, type(self).modulus))

    def __truediv__(self, other):
        o = self._coerce(other)
        if o is None:
            return NotImplemented
        return type(self)(self.value * pow(o, -1, type(self).modulus))

    def __rtruediv__(self, other):
        o = self._coerce(other)
        if o is None:
            return NotImplemented
        return type(self)(o * pow(self.value, -1, type(self).modulus))

    def __pow__(self, e):
        return type(self)(pow(self.value, int(e), type(self).modulus))

    def signed_(self):
        v = self.value
        return v - type(self).modulus if v > type(self).modulus >> 1 else v

    def unsigned_(self):
        return self.value


@functools.lru_cache(maxsize=None)
def _prime_field(modulus):
    name = f"GF({modulus})"
    cls = type(name, (PrimeFieldElement,), {"__slots__": ()})
    cls.modulus = modulus
    cls.order = modulus
    cls.characteristic = modulus
    cls.ext_deg = 1
    cls.is_signed = True
    return cls


def GF(modulus=None, **kwargs):
    """GF(p) for an int p, or GF(p^d) for an ExtPoly modulus (see extfield.py)."""
    if isinstance(modulus, int):
        return _prime_field(modulus)
    from . import extfield

    return extfield.ext_field(modulus)
