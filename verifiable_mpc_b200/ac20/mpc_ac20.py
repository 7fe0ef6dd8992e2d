"""Share-local group arithmetic of the MPC prover (``verifiable_mpc/ac20/mpc_ac20.py``), one GPU context per party.

In the reference every commitment of the multi-party prover goes through
``mpc_ac20.vector_commitment(x, gamma, g, h)`` (mpc_ac20.py:35-42) = ``secure_repeat(g + [h], x + [gamma])`` with
``secure_repeat = mpyc.secgroups.repeat_public_base_public_output`` (mpc_ac20.py:12): the bases are public, the
exponents are Shamir-shared, the output is public.  Each party i therefore computes LOCALLY

    c_i = prod_j  base_j ** (lambda_i * share_i(x_j))          (lambda_i: its Lagrange coefficient at 0)

sends c_i to the others, and everyone multiplies the m group elements together.  The local product is an (n+1)-term
multi-scalar multiplication -- exactly ``pivot.vector_commitment`` with the scaled shares as exponents -- and is what
this module runs on the party's own B200 (``VMSM_DEVICE`` / ``LOCAL_RANK`` selects it, ``fingroups._ctx``).  The share
vector is uploaded once; the scaling by lambda_i happens on the device (``vmsm_scalars_axpy``, SCALE) and the MSM reads
it in place.  The exchange of the m group elements and everything else secret-shared stay MPyC's
(INTEGRATION.md shows the ``secure_repeat`` rebinding).  ``recombine_at_zero`` is the m-point Lagrange vector of
``verifiable_mpc/ac20/recombine.py:5-37`` restricted to what this path needs.
"""
from .. import _lib
from ..engine import ED_L
from . import pivot


def recombine_at_zero(modulus, xs):
    """Lagrange coefficients lambda_i with sum_i lambda_i * f(xs[i]) = f(0) for deg f < len(xs) (recombine.py:5-37)."""
    xs = [x % modulus for x in xs]
    lam = []
    for i, x_i in enumerate(xs):
        num, den = 1, 1
        for j, x_j in enumerate(xs):
            if i != j:
                num = num * (0 - x_j) % modulus
                den = den * (x_i - x_j) % modulus
        lam.append(num * pow(den, -1, modulus) % modulus)
    return lam


def local_commitment_share(shares, gamma_share, g, h, lam):
    """Party-local factor of ``mpc_ac20.vector_commitment``: ``h**(lam*gamma_share) * prod g[j]**(lam*shares[j])``.

    ``shares`` / ``gamma_share``: this party's Shamir shares (ints or field elements); ``g``: DevicePointList or list
    of group elements; ``lam``: this party's recombination coefficient.  One device MSM.
    """
    assert len(g) >= len(shares), "Not enough generators."
    group = type(h)
    dev = pivot.as_device_list(g, group)
    ctx = dev.dev.ctx
    order = group.order
    hd = pivot._device_single(group, h)
    vals = [pivot._int(v) for v in shares]
    if order != ED_L:  # BN256 groups: scale on the host, one MSM with the blinding base as the extra term
        scaled = [v * lam % order for v in vals] + [pivot._int(gamma_share) * lam % order]
        return group._make(ctx.msm_ext(dev.dev, dev.off, len(vals), hd, 0, 1,
                                       pivot.pack_scalars(scaled, order)))
    sd = ctx.upload_scalars(vals, order)
    try:
        sd.axpy(lam, None, _lib.AXPY_SCALE)
        ctx.msm_dev_ext(dev.dev, dev.off, len(vals), sd, 0, hd, 0, [pivot._int(gamma_share) * lam % order], slot=0)
        return group._make(ctx.result(0))
    finally:
        sd.free()


def combine_commitment_shares(parts):
    """Product of the parties' local factors (what every party computes after the exchange)."""
    return pivot.list_mul(list(parts))


def vector_commitment_from_shares(share_rows, gamma_shares, g, h, xs=None):
    """All parties' work in one process (tests, single-host simulation): ``share_rows[i]`` are party i's shares of the
    exponent vector at evaluation point ``xs[i]`` (default i + 1, as MPyC numbers its parties)."""
    m = len(share_rows)
    group = type(h)
    xs = list(range(1, m + 1)) if xs is None else xs
    lam = recombine_at_zero(group.order, xs)
    return combine_commitment_shares(
        [local_commitment_share(share_rows[i], gamma_shares[i], g, h, lam[i]) for i in range(m)])
