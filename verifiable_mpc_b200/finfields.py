"""Minimal host-side prime fields with the surface of ``mpyc.finfields.GF`` that the AC20 prover touches
(SURVEY.md App. B.2): ``GF(modulus)`` -> class with ``.order`` / ``.modulus`` / ``.is_signed``, elements with
``+ - * / **``, ``int(e)`` (signed representative when ``is_signed``), ``.value``, ``== int``.

This is host scalar-field bookkeeping (challenges, linear forms), exactly what the reference keeps in Python; group
arithmetic never happens here.  When the real MPyC is installed its field classes work just as well -- the prover
modules only use the duck-typed surface above.
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
            return other.value if type(other) is type(self) else None
        if isinstance(other, int):
            return other
        return None

    def __eq__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else self.value == o % type(self).modulus

    def __add__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else type(self)(self.value + o)

    __radd__ = __add__

    def __sub__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else type(self)(self.value - o)

    def __rsub__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else type(self)(o - self.value)

    def __mul__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else type(self)(self.value * o)

    __rmul__ = __mul__

    def __neg__(self):
        return type(self)(-self.value)

    def reciprocal(self):
        return type(self)(pow(self.value, -1, type(self).modulus))

    def __truediv__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else type(self)(self.value * pow(o, -1, type(self).modulus))

    def __rtruediv__(self, other):
        o = self._coerce(other)
        return NotImplemented if o is None else type(self)(o * pow(self.value, -1, type(self).modulus))

    def __pow__(self, e):
        return type(self)(pow(self.value, int(e), type(self).modulus))


@functools.lru_cache(maxsize=None)
def GF(modulus):
    if not isinstance(modulus, int):
        raise NotImplementedError("only prime fields live in verifiable_mpc_b200.finfields")
    cls = type(f"GF({modulus})", (PrimeFieldElement,), {"__slots__": ()})
    cls.modulus = cls.order = cls.characteristic = modulus
    cls.ext_deg = 1
    cls.is_signed = True
    return cls
