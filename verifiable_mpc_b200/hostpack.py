"""Packing of the host-side vectors the reference's API hands over (lists of ints / field elements) into n x 32
little-endian residues.  Uses the C loop of ``_hostpack`` (csrc/_hostpack.c, built in-tree by
``__graft_entry__.build()``) when it is there, the same conversion in Python otherwise -- marshalling only, no group
or field arithmetic happens here."""
try:
    from . import _hostpack as _c
except ImportError:  # not built: pure Python, same results
    _c = None


def pack_residues(seq, cls, order, allow_int=True):
    """``seq``: exact ints (unless ``allow_int`` is False) and / or instances of exactly ``cls`` (a prime-field element
    class with modulus ``order`` and the residue in ``.value``; None: ints only).  -> bytes, or None when an element
    is of another type."""
    if _c is not None:
        return _c.pack_residues(seq, cls, order, allow_int)
    out = []
    for item in seq:
        t = type(item)
        if allow_int and t is int:
            v = item
        elif cls is not None and t is cls:
            v = item.value
            if not isinstance(v, int):
                return None
        else:
            return None
        if not 0 <= v < order:
            v %= order
        out.append(v.to_bytes(32, "little"))
    return b"".join(out)


def pack_auto(seq, order):
    """``pack_residues`` with the element class taken from the first non-int element (``field_class_for``'s rule) in
    the same pass: ints and / or elements of ONE prime-field class with this modulus -> bytes, anything else -> None."""
    if _c is not None:
        return _c.pack_residues(seq, True, order, True)
    cls = field_class_for(seq, order)
    return None if cls is False else pack_residues(seq, cls, order)


def field_class_for(seq, order):
    """The field-element class to expect in ``seq``: the type of its first non-int element when that is a prime-field
    class with this modulus, else None (ints only)."""
    for item in seq:
        if type(item) is not int:
            t = type(item)
            return t if getattr(t, "modulus", None) == order and hasattr(item, "value") else False
    return None


def mt_randbelow_packed(rng, order, n):
    """``[rng.randrange(order) for _ in range(n)]`` of a seeded ``random.Random`` as n x 32 little-endian bytes, leaving
    ``rng`` in exactly the state the n calls would have left it (C loop over the generator's own Mersenne-twister state,
    csrc/_hostpack.c); None when the helper is not built or ``rng`` is not a plain ``random.Random``."""
    import random as _random
    import struct

    if _c is None or type(rng) is not _random.Random or not hasattr(_c, "mt_randbelow_packed") or order <= 1 or order >> 256:
        return None
    version, state, gauss = rng.getstate()
    if version != 3 or len(state) != 625:
        return None
    out, key, pos = _c.mt_randbelow_packed(struct.pack("<624I", *state[:624]), state[624], order, n)
    rng.setstate((version, struct.unpack("<624I", key) + (pos,), gauss))
    return out
