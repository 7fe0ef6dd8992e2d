"""Affine / linear forms over the scalar field: the host-side bookkeeping objects of the AC20 protocols.

Behaviourally equal to ``AffineForm`` / ``LinearForm`` of verifiable_mpc/ac20/pivot.py:31-116 (same operators, same
``repr`` -- which enters the Fiat-Shamir pre-image -- same assertion texts, sums of linear forms fall back to affine
forms); re-exported by ``verifiable_mpc_b200.ac20.pivot`` under the reference's names.
"""
from ..finfields import FiniteFieldElement as _OwnFieldElement

import sys


class SecureObject:
    """Placeholder exported under the reference's name; real MPyC secure objects are recognised by `secure_types()`."""


def field_types():
    """Field-element classes accepted as scalars: this package's, and MPyC's when `mpyc.finfields` is loaded -- looked
    up at call time, so it does not matter whether MPyC (or a look-alike) was imported before or after this package."""
    mod = sys.modules.get("mpyc.finfields")
    t = getattr(mod, "FiniteFieldElement", None) if mod is not None else None
    return (_OwnFieldElement, t) if isinstance(t, type) else (_OwnFieldElement,)


def secure_types():
    mod = sys.modules.get("mpyc.sectypes")
    t = getattr(mod, "SecureObject", None) if mod is not None else None
    return (SecureObject, t) if isinstance(t, type) else (SecureObject,)


def _is_scalar(v):
    return isinstance(v, (int,) + secure_types() + field_types())


class AffineForm:
    """f(x) = <coeffs, x> + constant over the scalar field (host-side bookkeeping, as in the reference)."""

    def __init__(self, coeffs, constant):
        self.coeffs = coeffs
        self.constant = constant

    def _combine(self, other):
        if isinstance(other, AffineForm):
            assert len(self) == len(other), "Length of linear forms to add not consistent."
            return [a + b for a, b in zip(self.coeffs, other.coeffs)], self.constant + other.constant
        if _is_scalar(other):
            return self.coeffs, self.constant + other
        return None

    def __add__(self, other):
        res = self._combine(other)
        if res is None:
            raise NotImplementedError(f"Addition of form not defined for type: {type(other)}")
        return type(self)(*res)

    def __radd__(self, other):
        return self if other == 0 else self.__add__(other)

    def __sub__(self, other):
        return self + (-1) * other

    def __mul__(self, other):
        if not isinstance(other, (int,) + field_types()):
            raise NotImplementedError(f"Multiplication of form not defined for type: {type(other)}")
        return type(self)([c * other for c in self.coeffs], self.constant * other)

    __rmul__ = __mul__

    def __len__(self):
        return len(self.coeffs)

    def __eq__(self, other):
        return self.coeffs == other.coeffs

    def __repr__(self):
        return f"{str(self.coeffs)}, {str(self.constant)}"

    def eval(self, values):
        assert len(values) == len(self.coeffs), "Length of inputs to be equal to coefficients of linear form."
        return sum([c * v for c, v in zip(self.coeffs, values)]) + self.constant

    __call__ = eval


class LinearForm(AffineForm):
    """Affine form whose constant is pinned to 0; sums fall back to AffineForm like the reference's."""

    def __init__(self, coeffs, constant=0):
        super().__init__(coeffs, 0)

    def __add__(self, other):
        res = self._combine(other)
        if res is None:
            raise NotImplementedError
        return AffineForm(*res)
