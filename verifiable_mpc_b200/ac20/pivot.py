"""GPU-backed twin of ``verifiable_mpc/ac20/pivot.py`` (AC20 protocol 2, "pivot").

Same names, argument meaning, return values and error behaviour as the reference module so callers can rebind it
(SURVEY.md 8b1):  ``vector_commitment`` (pivot.py:139-145), ``list_mul`` (:26-28), ``fiat_shamir_hash`` (:131-136),
``AffineForm`` / ``LinearForm`` (:31-116), ``_int`` (:119-128), ``affine_to_linear`` (:148-153),
``prove_linear_form_eval`` (:156-181), ``verify_linear_form_proof`` (:184-205).

What changes: the n independent ``g[i] ** x_i`` scalar multiplications + tree product become ONE Pippenger MSM on
the device (libvmsm.so: vmsm_msm_ext), generators may stay resident in HBM (``DevicePointList``), and every other
group operation is a device call too.  Linear-form algebra and hashing stay on the host exactly as in the reference.
"""
import hashlib
import logging
from random import SystemRandom

from .. import fingroups
from ..engine import pack_scalars
from ..fingroups import DevicePointList, EllipticCurvePoint as EllipticCurveElement

from .. import hostpack
from .forms import field_types as _field_types, secure_types as _secure_types
from .forms import AffineForm, LinearForm, SecureObject  # noqa: F401  (the reference exports them from pivot)

prng = SystemRandom()

logger_piv = logging.getLogger("pivot")
logger_piv.setLevel(logging.INFO)


def random_residues(rng, order, n):
    """``[rng.randrange(order) for _ in range(n)]``, same draws from the same generator state.  For the stdlib
    generators (random.Random / SystemRandom, what the reference uses: pivot.py:21, compressed_pivot.py:22) the
    rejection loop of ``Random._randbelow_with_getrandbits`` is run here without three Python calls per draw."""
    import random as _random

    if type(rng) in (_random.Random, _random.SystemRandom) and order > 0:
        bits = order.bit_length()
        getrandbits = rng.getrandbits
        out = []
        append = out.append
        for _ in range(n):
            v = getrandbits(bits)
            while v >= order:
                v = getrandbits(bits)
            append(v)
        return out
    return [rng.randrange(order) for _ in range(n)]


def random_residues_packed(rng, order, n):
    """The draws of ``random_residues(rng, order, n)`` as n x 32 little-endian bytes (what the device wants), leaving
    the generator in the same state, without creating n Python ints; None when ``rng`` is not a stdlib generator.

    ``Random.randrange(order)`` is ``getrandbits(bits)`` repeated until the value is below ``order``, and
    ``getrandbits(bits)`` consumes ceil(bits / 32) words of the Mersenne twister, least significant first, the last
    one shifted down to the remaining bits.  One ``getrandbits`` of many words therefore yields the same words in
    the same order; numpy applies the shift and the rejection test to all candidates of a batch at once."""
    import random as _random

    import numpy as np

    bits = order.bit_length()
    if type(rng) not in (_random.Random, _random.SystemRandom) or order <= 1 or bits > 256 or n == 0:
        return None
    if type(rng) is _random.Random:  # the same stream from a C loop over the generator's own state (hostpack)
        raw = hostpack.mt_randbelow_packed(rng, order, n)
        if raw is not None:
            return np.frombuffer(raw, dtype=np.uint8).reshape(n, 32)
    words = (bits + 31) // 32
    shift = 32 * words - bits
    limbs = [(order >> (32 * k)) & 0xFFFFFFFF for k in range(words)]
    need, chunks = n, []
    while need:
        # `need` more accepted values take at least `need` more candidates: drawing exactly that many never consumes
        # a word the sequential rejection loop would not have consumed, so the generator ends in the same state
        m = need
        raw = rng.getrandbits(32 * words * m).to_bytes(4 * words * m, "little")
        cand = np.frombuffer(raw, dtype="<u4").reshape(m, words)
        if shift:
            cand = cand.copy()
            cand[:, -1] >>= shift
        below = np.zeros(m, dtype=bool)
        for k in range(words):  # lexicographic candidate < order, least significant limb first
            below = (cand[:, k] < limbs[k]) | ((cand[:, k] == limbs[k]) & below)
        idx = np.flatnonzero(below)
        chunks.append(cand[idx])
        need -= len(idx)
    out = np.zeros((n, 8), dtype="<u4")
    out[:, :words] = np.concatenate(chunks) if len(chunks) > 1 else chunks[0]
    return out.view(np.uint8).reshape(n, 32)


def _int(value):
    """Field elements -> ints (signed representative, as MPyC's int()); ints and secure objects pass through."""
    if isinstance(value, (int,) + _secure_types()):
        return value
    if isinstance(value, _field_types()):
        return int(value)
    raise NotImplementedError


def fiat_shamir_hash(input_list, order):
    digest = hashlib.sha256(str(input_list).encode("utf-8")).digest()
    return int.from_bytes(digest, "little") % order


# ---------------------------------------------------------------------------------------------- binary transcript
# Opt-in alternative to the reference's pre-images (SHA-256 over str(list): ~157 characters of decimal text per
# generator per round).  TRANSCRIPT = "binary" hashes canonical bytes instead -- 64 bytes x || y per group element,
# 32 bytes per scalar, every item tagged and length-prefixed, one domain-separation label per call site -- which
# removes the text formatting and 60 % of the hashing.  It changes every challenge, so prover and verifier must both
# run this package with the same setting; proofs made in this mode are NOT verifiable by the reference.
TRANSCRIPT = "reference"
_BIN_DOMAIN = b"verifiable_mpc_b200 transcript v1\x00"


def _feed_binary(h, item, order):
    """One transcript item: tag byte, 8-byte little-endian count, canonical bytes."""
    wire = getattr(item, "wire_bytes", None)
    if wire is not None:  # device-resident generator vector
        h.update(b"V" + len(item).to_bytes(8, "little"))
        h.update(wire())
    elif hasattr(item, "affine"):  # group element
        x, y = item.affine()
        h.update(b"P" + int(x).to_bytes(32, "little") + int(y).to_bytes(32, "little"))
    elif hasattr(item, "transcript_scalars"):  # linear / affine form stand-ins of compressed_pivot
        n, coeff_bytes, constant = item.transcript_scalars()
        h.update(b"F" + n.to_bytes(8, "little"))
        h.update(coeff_bytes)
        h.update((_int(constant) % order).to_bytes(32, "little"))
    elif isinstance(item, AffineForm):
        h.update(b"F" + len(item.coeffs).to_bytes(8, "little"))
        h.update(pack_scalars([_int(v) for v in item.coeffs], order))
        h.update((_int(item.constant) % order).to_bytes(32, "little"))
    elif type(item) is dict:  # the generators argument
        h.update(b"D" + len(item).to_bytes(8, "little"))
        for key, value in item.items():
            h.update(str(key).encode("utf-8") + b"\x00")
            _feed_binary(h, value, order)
    elif isinstance(item, (list, tuple)):
        if item and all(hasattr(v, "affine") for v in item):  # a host list of generators: same bytes as a device vector
            h.update(b"V" + len(item).to_bytes(8, "little"))
            h.update(b"".join(int(c).to_bytes(32, "little") for v in item for c in v.affine()))
            return
        h.update(b"L" + len(item).to_bytes(8, "little"))
        for value in item:
            _feed_binary(h, value, order)
    elif isinstance(item, (bytes, str)):
        raw = item.encode("utf-8") if isinstance(item, str) else item
        h.update(b"T" + len(raw).to_bytes(8, "little") + raw)
    else:  # scalar: int or field element
        h.update(b"S" + (_int(item) % order).to_bytes(32, "little"))


def binary_prefix(label, items, order):
    h = hashlib.sha256()
    h.update(_BIN_DOMAIN + label + b"\x00")
    for item in items:
        _feed_binary(h, item, order)
    return h


def binary_finish(h, tail_items, order):
    h = h.copy()
    for item in tail_items:
        _feed_binary(h, item, order)
    return int.from_bytes(h.digest(), "little") % order


def transcript_challenge(label, items, order):
    """The challenge for one call site: the reference's str(list) pre-image, or the binary encoding when opted in."""
    if TRANSCRIPT == "binary":
        return binary_finish(binary_prefix(label, items, order), [], order)
    return fiat_shamir_hash_items(items, order)


def _feed_repr(h, item):
    """h.update(repr(item).encode()) without building the large strings: device-resident generator lists offer
    ``repr_bytes()`` (decimal text produced on the GPU); dicts (the ``generators`` argument) are walked."""
    feed = getattr(item, "feed_repr", None)
    rb = getattr(item, "repr_bytes", None)
    if feed is not None:
        feed(h)
    elif rb is not None:
        h.update(rb())
    elif type(item) is dict:
        h.update(b"{")
        for i, (key, value) in enumerate(item.items()):
            h.update((", " if i else "").encode() + repr(key).encode("utf-8") + b": ")
            _feed_repr(h, value)
        h.update(b"}")
    else:
        h.update(repr(item).encode("utf-8"))


def fiat_shamir_prefix(items):
    """SHA-256 state after ``"[" + ", ".join(repr(i) for i in items)`` -- the shared part of str(list) pre-images."""
    h = hashlib.sha256()
    h.update(b"[")
    for i, item in enumerate(items):
        if i:
            h.update(b", ")
        _feed_repr(h, item)
    return h


def fiat_shamir_finish(h, tail_items, order):
    """Challenge from a prefix state plus the remaining list items (the state is copied, so it can be reused)."""
    h = h.copy()
    for item in tail_items:
        h.update(b", ")
        _feed_repr(h, item)
    h.update(b"]")
    return int.from_bytes(h.digest(), "little") % order


def fiat_shamir_hash_items(items, order):
    """Same value as ``fiat_shamir_hash(items, order)``, computed incrementally (no O(N) Python str)."""
    return fiat_shamir_finish(fiat_shamir_prefix(items), [], order)


def list_mul(x):
    """Product of a list of group elements: one device call per 64 elements instead of len(x)-1 host operations."""
    rettype = type(x[0])
    acc = rettype.identity
    for i in range(0, len(x), 63):
        chunk = [acc] + list(x[i:i + 63])
        acc = rettype.lincomb(chunk, [1] * len(chunk))
    return acc


# ---------------------------------------------------------------------------------------------- device plumbing
_single_cache = {}


def _device_single(group, pt):
    """One-element device vector for a blinding base (h or k), cached per context."""
    ctx = group._ctx()
    key = (id(ctx), pt.affine())
    hit = _single_cache.get(key)
    if hit is None or not hit.handle:
        if len(_single_cache) > 64:
            _single_cache.clear()
        hit = ctx.upload_points([pt.affine()])
        _single_cache[key] = hit
    return hit


def as_device_list(g, group=None):
    """Generator list -> DevicePointList (uploads a Python list of points once; passes device lists through)."""
    if isinstance(g, DevicePointList):
        return g
    group = group or type(g[0])
    return DevicePointList.from_points(group, g)


def vector_commitment(x, gamma, g, h):
    """Pedersen vector commitment ``h**gamma * prod g[i]**x[i]`` (AC20 definition 1) as ONE device MSM.

    ``g``: DevicePointList or list of group elements; ``x`` / ``gamma``: ints or field elements, negative and
    unreduced values allowed exactly as the reference passes them.
    """
    assert len(g) >= len(x), "Not enough generators."
    group = type(h)
    dev = as_device_list(g, group)
    order = group.order
    # ints and elements of the exponent field in one C pass (hostpack); anything else through the reference's _int
    raw = hostpack.pack_auto(x, order) if isinstance(x, (list, tuple)) else None
    if raw is None:
        raw = pack_scalars([_int(v) for v in x], order)
    raw += pack_scalars([_int(gamma)], order)
    hd = _device_single(group, h)
    xy = dev.dev.ctx.msm_ext(dev.dev, dev.off, len(x), hd, 0, 1, raw)
    return group._make(xy)


def affine_to_linear(L, y, n):
    constant = L([0] * n)
    return L - constant, y - constant


def prove_linear_form_eval(g, h, P, L, y, x, gamma, gf):
    """Sigma protocol Pi_s, non-interactive (reference pivot.py:156-181)."""
    n = len(x)
    L, y = affine_to_linear(L, y, n)
    r = [gf(prng.randrange(gf.order)) for _ in range(n)]
    rho = prng.randrange(gf.order)
    t = L(r)
    A = vector_commitment(r, rho, g, h)
    logger_piv.debug(f"Prover computed A={A}.")
    c = transcript_challenge(b"pivot", [t, A.normalize(), g, h, P.normalize(), L, y], gf.order)
    z = [c * x_i + r_i for x_i, r_i in zip(x, r)]
    phi = (c * gamma + rho) % gf.order
    return z, phi, c


def verify_linear_form_proof(g, h, P, L, y, z, phi, c):
    n = len(z)
    L, y = affine_to_linear(L, y, n)
    group = type(P)
    A_check = group.lincomb([vector_commitment(z, phi, g, h), P], [1, -int(c)])
    t_check = L(z) - c * y
    order = type(t_check).order
    hash_check = transcript_challenge(b"pivot", [t_check, A_check.normalize(), g, h, P.normalize(), L, y], order)
    logger_piv.debug(f"Value of c         ={c}")
    logger_piv.debug(f"Value of hash_check={hash_check}")
    return c == hash_check
