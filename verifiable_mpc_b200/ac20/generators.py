"""Batch generator creation on the device: the compressed-pivot branch of ``create_generators``
(verifiable_mpc/ac20/circuit_sat_r1cs.py:47-93: ``g_i = h ** r_i`` for random r_i, ``k = h ** r``, h = group.generator)
as ONE fixed-base kernel launch instead of g_length Python scalar multiplications.  The generators stay in HBM."""
from random import SystemRandom

from ..engine import pack_scalars
from ..fingroups import DevicePointList

prng = SystemRandom()


def create_generators(g_length, group, with_k=True, exponents=None):
    """-> {"g": DevicePointList(g_length), "h": group.generator, "k": point}  (``with_k=False``: basic pivot)."""
    if getattr(group, "curve_id", 0) != 0:
        # DevicePointList and the fixed-base launch below are the Ed25519 wire format; the reference's loop works for
        # any group, so say so instead of returning Ed25519 points tagged as another group
        raise NotImplementedError("create_generators on the device is implemented for the Ed25519 group only")
    ctx = group._ctx()
    if exponents is None:
        exponents = [prng.randrange(1, group.order) for _ in range(g_length)]
    assert len(exponents) == g_length
    dev = ctx.fixed_base(scalars=pack_scalars(exponents, group.order))
    generators = {"g": DevicePointList(group, dev), "h": group.generator}
    if with_k:
        generators["k"] = group.generator ** prng.randrange(1, group.order) if group.is_multiplicative else \
            group.repeat(group.generator, prng.randrange(1, group.order))
    return generators
