"""GPU-backed twin of ``verifiable_mpc/ac20/compressed_pivot.py`` (AC20 protocols 4 and 5, compressed pivot).

Same signatures, proof-dict keys ("t", "A", "A{i}", "B{i}", "z_prime") and Fiat-Shamir pre-images as the reference
(compressed_pivot.py:29-239), so a proof made here verifies with the reference's verifier running on
group types that print canonical affine coordinates, and vice versa.  Per folding round the device does: two
(half+1)-term MSMs (A_i, B_i), one 3-term combination (Q'), and ONE fold kernel over the generator vector kept in
HBM (``g'_j = c * g_j + g_{half+j}``, compressed_pivot.py:64 / :178), and -- above DEVICE_SCALAR_MIN entries -- the
halving of the witness and of the linear form, their cross terms and their decimal text (csrc/sc25519.cuh).  The host
draws the randomness, hashes the transcript bytes the device formats, and finishes the last few rounds on integers.
"""
import logging
from concurrent.futures import ThreadPoolExecutor
from random import SystemRandom

from . import pivot
from .. import _lib, hostpack
from ..engine import ED_L
from ..fingroups import DevicePointList

prng = SystemRandom()


FIRST_HASH_THREAD_MIN = 4096


def _WORKER():
    return ThreadPoolExecutor(max_workers=1)


logger_cp = logging.getLogger("compressed_pivot")
logger_cp.setLevel(logging.INFO)
logger_cp_hin = logging.getLogger("compressed_pivot_hash_inputs")
logger_cp_hin.setLevel(logging.INFO)
logger_cp_hout = logging.getLogger("compressed_pivot_hash_outputs")
logger_cp_hout.setLevel(logging.INFO)

_TAG = "First hash of compressed pivot"

# Host scalar algebra.  The reference keeps linear forms and witnesses as lists of field-element OBJECTS and pays one
# Python method call per coefficient per round; when every entry is an element of `gf` (the normal case) the twins
# below run the same algebra on plain ints modulo the order and only materialise objects / text where the reference's
# outputs require it (proof["z_prime"], the decimal text inside the Fiat-Shamir pre-image).  Mixed / plain-int inputs
# take the generic object path, which reproduces the reference's unreduced-integer behaviour exactly.
FAST_INT_PATH = True


def _all_in_field(values, gf):
    return set(map(type, values)) <= {gf}


def _pack_field(values, gf, order):
    """n x 32 bytes of the residues of a list of gf elements (hostpack: one C loop when the helper module is built)."""
    raw = hostpack.pack_residues(values, gf, order, False)
    if raw is None:  # a look-alike element class: through .value in Python
        raw = pivot.pack_scalars([v.value for v in values], order)
    return raw


def _field_text(ints, q, signed=True):
    """repr of a list of gf elements from their residues: signed representatives when the field class is signed
    (MPyC's default and what int(element) returns), plain residues otherwise."""
    if not signed:
        return "[" + ", ".join([str(v) for v in ints]) + "]"
    half = q >> 1
    return "[" + ", ".join([str(v - q if v > half else v) for v in ints]) + "]"


class _IntForm:
    """Stand-in for a LinearForm over gf inside the round loop: residues + the exact repr of the real thing."""
    __slots__ = ("ints", "q", "signed")

    def __init__(self, ints, q, signed=True):
        self.ints, self.q, self.signed = ints, q, signed

    def __repr__(self):
        return f"{_field_text(self.ints, self.q, self.signed)}, 0"

    def transcript_scalars(self):
        return len(self.ints), pivot.pack_scalars(self.ints, self.q), 0


def _dot(a, b, q):
    return sum(map(int.__mul__, a, b)) % q


def _private_device_list(g_hat, group):
    """A device vector this call may fold in place (the caller's list stays untouched, like the reference's)."""
    dev = pivot.as_device_list(g_hat, group)
    if getattr(dev, "_owned", False):
        return dev
    own = dev.clone()
    own._owned = True
    return own


_Q_SLOT = 5  # result slot of the asynchronous Q' (the A_i / B_i commitments of the device loop use slots 0 and 1)


class _PendingPoint:
    """Q' = A * Q**c * B**(c**2) issued asynchronously: it is awaited by the next challenge hash (or the final
    check), so it is computed underneath the generator / witness folds and the next round's commitments."""
    __slots__ = ("group", "ctx")

    def __init__(self, group, ctx):
        self.group, self.ctx = group, ctx

    def get(self):
        return self.group._make(self.ctx.result(_Q_SLOT))


def _q_prime(group, A, Q, B, c):
    ctx = group._ctx()
    if getattr(group, "curve_id", 0) == 0 and hasattr(ctx, "lincomb_async"):
        order = group.order
        ctx.lincomb_async([A.affine(), Q.affine(), B.affine()], [1, c % order, c * c % order], _Q_SLOT)
        return _PendingPoint(group, ctx)
    return group.lincomb([A, Q, B], [1, c, c ** 2])


def _resolve(Q):
    return Q.get() if type(Q) is _PendingPoint else Q


def _fold_challenge(A, B, g_hat, k, Q, L_tilde, order):
    input_list = [A.normalize(), B.normalize(), g_hat, k, Q.normalize(), L_tilde]
    if logger_cp_hin.isEnabledFor(logging.DEBUG):
        logger_cp_hin.debug(f"Before fiat_shamir_hash, input_list=\n{input_list}")
    c = pivot.transcript_challenge(b"cp-round", input_list, order)
    logger_cp_hout.debug(f"After hash, hash=\n{c}")
    return c


def _fold_forms(L_tilde, c, half, gf):
    assert L_tilde.constant == 0, "Next line assumes L_tilde is a linear form, not affine form."
    left = [coeff * gf(c) for coeff in L_tilde.coeffs[:half]]
    return pivot.LinearForm(left) + pivot.LinearForm(L_tilde.coeffs[half:])


def _fold_generators(g_hat, c):
    """In place on the device: first half <- c * first half + second half; the view shrinks to `half`."""
    half = len(g_hat) // 2
    assert g_hat.off == 0 and g_hat.n == g_hat.dev.n
    g_hat.dev.fold(c)
    g_hat.n = half
    return g_hat


def protocol_4_prover(g_hat, k, Q, L_tilde, z_hat, gf, proof=None, round_i=0):
    """Non-interactive protocol 4, prover (reference :29-86).  Iterative instead of recursive; same rounds."""
    proof = {} if proof is None else proof
    group = type(k)
    g_hat = _private_device_list(g_hat, group)
    order = k.order
    if (FAST_INT_PATH and isinstance(L_tilde, pivot.LinearForm) and gf.order == order
            and _all_in_field(L_tilde.coeffs, gf) and _all_in_field(z_hat, gf)):
        return _protocol_4_prover_fast(g_hat, k, Q, [c.value for c in L_tilde.coeffs], [z.value for z in z_hat], gf,
                                       proof, round_i)
    while True:
        half = len(g_hat) // 2
        z_l, z_r = z_hat[:half], z_hat[half:]
        logger_cp.debug("Calculate A_i, B_i.")
        A = pivot.vector_commitment(z_l, int(L_tilde([0] * half + z_l)), g_hat[half:], k)
        B = pivot.vector_commitment(z_r, int(L_tilde(z_r + [0] * half)), g_hat[:half], k)
        proof["A" + str(round_i)] = A
        proof["B" + str(round_i)] = B
        Q = _resolve(Q)
        c = _fold_challenge(A, B, g_hat, k, Q, L_tilde, order)
        logger_cp.debug("Calculate g_prime.")
        g_hat = _fold_generators(g_hat, c)
        logger_cp.debug("Calculate Q_prime.")
        Q = _q_prime(group, A, Q, B, c)
        L_tilde = _fold_forms(L_tilde, c, half, gf)
        z_hat = [l + c * r for l, r in zip(z_l, z_r)]
        if len(z_hat) <= 2:
            proof["z_prime"] = z_hat
            return proof
        round_i += 1


def _protocol_4_prover_ints(g_hat, k, Q, coeffs, z, gf, proof, round_i):
    """The round loop of protocol_4_prover on residues modulo the group order (same values, same transcript)."""
    group = type(k)
    q = k.order
    while True:
        half = len(g_hat) // 2
        z_l, z_r = z[:half], z[half:]
        logger_cp.debug("Calculate A_i, B_i.")
        A = pivot.vector_commitment(z_l, _dot(coeffs[half:], z_l, q), g_hat[half:], k)
        B = pivot.vector_commitment(z_r, _dot(coeffs[:half], z_r, q), g_hat[:half], k)
        proof["A" + str(round_i)] = A
        proof["B" + str(round_i)] = B
        Q = _resolve(Q)
        c = _fold_challenge(A, B, g_hat, k, Q, _IntForm(coeffs, q, bool(gf.is_signed)), q)
        g_hat = _fold_generators(g_hat, c)
        Q = _q_prime(group, A, Q, B, c)
        coeffs = [(l * c + r) % q for l, r in zip(coeffs[:half], coeffs[half:])]
        z = [(l + c * r) % q for l, r in zip(z_l, z_r)]
        if len(z) <= 2:
            proof["z_prime"] = [gf(v) for v in z]
            return proof
        round_i += 1


def _protocol_4_verifier_ints(g_hat, k, Q, coeffs, gf, proof, round_i):
    group = type(k)
    q = k.order
    while True:
        half = len(g_hat) // 2
        A = proof["A" + str(round_i)]
        B = proof["B" + str(round_i)]
        Q = _resolve(Q)
        c = _fold_challenge(A, B, g_hat, k, Q, _IntForm(coeffs, q, bool(gf.is_signed)), q)
        g_hat = _fold_generators(g_hat, c)
        Q = _q_prime(group, A, Q, B, c)
        coeffs = [(l * c + r) % q for l, r in zip(coeffs[:half], coeffs[half:])]
        if len(g_hat) <= 2:
            return _final_check(g_hat, k, Q, coeffs, gf, proof)
        round_i += 1


def _final_check(g_hat, k, Q, coeffs, gf, proof):
    Q = _resolve(Q)
    z_prime = proof["z_prime"]
    # the proof is untrusted input: the reference asserts on a length mismatch inside L_prime(z_prime) (pivot.py:62-64)
    assert len(z_prime) == len(coeffs) == len(g_hat), "Length of input vector and linear form do not match."
    if not all(isinstance(zp, (int, gf)) for zp in z_prime):
        raise NotImplementedError
    gamma = int(sum([gf(cf) * zp for cf, zp in zip(coeffs, z_prime)]) + 0)
    Q_check = pivot.vector_commitment(z_prime, gamma, g_hat, k)
    return Q_check == Q


# Device-resident witness and linear form.  Above DEVICE_SCALAR_MIN entries the round loop keeps z and the coefficients
# of L_tilde in HBM next to the generators: the cross terms L_R(z_L), L_L(z_R) are device dot products, A_i / B_i read
# z in place, the halvings z' = z_L + c z_R and L' = c L_L + L_R are one kernel each, and the decimal text of L_tilde
# for the next challenge is produced on the device like that of g_hat.  Per round the host only hashes.  Below the
# threshold (a few rounds, a few hundred scalars in total) the vectors come back and the integer loop finishes.
DEVICE_SCALAR_PATH = True
DEVICE_SCALAR_MIN = 256          # verifier (no commitments per round: the integer loop is as fast below this)
DEVICE_SCALAR_MIN_PROVER = 2     # prover: A_i and B_i are issued as an asynchronous pair only on the device path
FUSED_CROSS_TERMS = True         # cross terms feed their commitments on the device (vmsm_msm_dev_ext_dot)


class _DevForm:
    """The form whose coefficients live on the device, for the Fiat-Shamir pre-image ("[coeffs], constant")."""
    __slots__ = ("sc", "n", "signed", "constant", "view")

    def __init__(self, sc, n, signed, constant="0"):
        self.sc, self.n, self.signed, self.constant, self.view = sc, n, signed, str(constant), None

    def prefetch_repr(self):
        """Coefficient text fetched ahead of the hash (see DevicePointList.prefetch_repr); this object describes one
        state of the vector and is discarded with it."""
        if hasattr(self.sc, "text_view"):
            self.view = self.sc.text_view(0, self.n, self.signed)
        return self

    def repr_bytes(self):
        return b"[" + self.sc.text_bytes(0, self.n, self.signed) + b"], " + self.constant.encode("ascii")

    def feed_repr(self, h):
        if hasattr(self.sc, "text_view"):  # hashed in place from the pinned buffer
            view, self.view = self.view, None
            h.update(b"[")
            h.update(view if view is not None else self.sc.text_view(0, self.n, self.signed))
            h.update(b"], " + self.constant.encode("ascii"))
        else:
            h.update(self.repr_bytes())

    def transcript_scalars(self):
        view = self.sc.wire_view(0, self.n) if hasattr(self.sc, "wire_view") else self.sc.download(0, self.n)
        return self.n, view, int(self.constant)

    def __repr__(self):
        return self.repr_bytes().decode("ascii")


PREFETCH_TEXT = True  # device prover: fetch the round's transcript text underneath the A_i, B_i commitments

TRACE = None  # tools/profile_ac20.py sets a list: (label, perf_counter()) marks of the device-resident prover


def _mark(label):
    if TRACE is not None:
        import time

        TRACE.append((label, time.perf_counter()))


def _use_device_scalars(g_hat, q, min_n=None):
    n = len(g_hat)
    min_n = DEVICE_SCALAR_MIN if min_n is None else min_n
    return (DEVICE_SCALAR_PATH and isinstance(g_hat, DevicePointList) and n > min_n and n & (n - 1) == 0
            and q == ED_L)


def _protocol_4_prover_fast(g_hat, k, Q, coeffs, z, gf, proof, round_i):
    q = k.order
    if (not _use_device_scalars(g_hat, q, DEVICE_SCALAR_MIN_PROVER) or len(z) != len(g_hat)
            or len(coeffs) != len(g_hat)):
        return _protocol_4_prover_ints(g_hat, k, Q, coeffs, z, gf, proof, round_i)
    ctx = g_hat.dev.ctx
    return _protocol_4_prover_dev(g_hat, k, Q, ctx.upload_scalars(coeffs, q), ctx.upload_scalars(z, q), gf, proof, round_i)


def _protocol_4_prover_dev(g_hat, k, Q, Ld, zd, gf, proof, round_i):
    """Round loop with L_tilde (Ld) and z_hat (zd) resident on the device; takes ownership of both vectors."""
    group = type(k)
    q = k.order
    ctx = g_hat.dev.ctx
    signed = bool(gf.is_signed)
    kd = pivot._device_single(group, k)
    min_n = DEVICE_SCALAR_MIN_PROVER
    fused = FUSED_CROSS_TERMS and hasattr(ctx, "msm_dev_ext_dot")
    try:
        n = len(g_hat)
        while n > min_n:
            half = n // 2
            logger_cp.debug("Calculate A_i, B_i.")
            _mark("round:start")
            if fused:
                # the cross terms L_tilde([0]*half + z_L), L_tilde(z_R + [0]*half) stay on the device: each is the
                # scalar of the k-term of its commitment (no host round trip between the inner product and the MSM)
                ctx.msm_dev_ext_dot(g_hat.dev, g_hat.off + half, half, zd, 0, kd, 0, Ld, half, zd, 0, half, slot=0)
                ctx.msm_dev_ext_dot(g_hat.dev, g_hat.off, half, zd, half, kd, 0, Ld, 0, zd, half, half, slot=1)
            else:
                s_a = ctx.scalars_dot(Ld, half, zd, 0, half)  # L_tilde([0]*half + z_L)
                s_b = ctx.scalars_dot(Ld, 0, zd, half, half)  # L_tilde(z_R + [0]*half)
                _mark("round:dots")
                ctx.msm_dev_ext(g_hat.dev, g_hat.off + half, half, zd, 0, kd, 0, [s_a], slot=0)
                ctx.msm_dev_ext(g_hat.dev, g_hat.off, half, zd, half, kd, 0, [s_b], slot=1)
            # the O(n) text of this round's pre-image (generators, coefficients of L_tilde) does not depend on A_i, B_i:
            # it is produced and copied to the host while the two commitments are being computed
            form = _DevForm(Ld, n, signed)
            if PREFETCH_TEXT and pivot.TRANSCRIPT != "binary":
                g_hat.prefetch_repr()
                form.prefetch_repr()
                _mark("round:text")
            A, B = group._make(ctx.result(0)), group._make(ctx.result(1))
            _mark("round:A,B")
            proof["A" + str(round_i)] = A
            proof["B" + str(round_i)] = B
            Q = _resolve(Q)
            c = _fold_challenge(A, B, g_hat, k, Q, form, q)
            _mark("round:challenge")
            g_hat = _fold_generators(g_hat, c)
            Q = _q_prime(group, A, Q, B, c)
            Ld.fold(half, c, _lib.FOLD_FORM)
            zd.fold(half, c, _lib.FOLD_WITNESS)
            _mark("round:folds issued")
            n = half
            if n <= 2:
                proof["z_prime"] = [gf(v) for v in zd.tolist(0, n)]
                return proof
            round_i += 1
        coeffs, z = Ld.tolist(0, n), zd.tolist(0, n)
    finally:
        zd.free()
        Ld.free()
    return _protocol_4_prover_ints(g_hat, k, Q, coeffs, z, gf, proof, round_i)


def _protocol_4_verifier_fast(g_hat, k, Q, coeffs, gf, proof, round_i):
    q = k.order
    if not _use_device_scalars(g_hat, q) or len(coeffs) != len(g_hat):
        return _protocol_4_verifier_ints(g_hat, k, Q, coeffs, gf, proof, round_i)
    return _protocol_4_verifier_dev(g_hat, k, Q, g_hat.dev.ctx.upload_scalars(coeffs, q), gf, proof, round_i)


def _protocol_4_verifier_dev(g_hat, k, Q, Ld, gf, proof, round_i):
    """Verifier rounds with L_tilde resident on the device; takes ownership of Ld."""
    group = type(k)
    q = k.order
    signed = bool(gf.is_signed)
    try:
        n = len(g_hat)
        while n > DEVICE_SCALAR_MIN:
            half = n // 2
            A = proof["A" + str(round_i)]
            B = proof["B" + str(round_i)]
            Q = _resolve(Q)
            c = _fold_challenge(A, B, g_hat, k, Q, _DevForm(Ld, n, signed), q)
            g_hat = _fold_generators(g_hat, c)
            Q = _q_prime(group, A, Q, B, c)
            Ld.fold(half, c, _lib.FOLD_FORM)
            n = half
            if n <= 2:
                return _final_check(g_hat, k, Q, Ld.tolist(0, n), gf, proof)
            round_i += 1
        coeffs = Ld.tolist(0, n)
    finally:
        Ld.free()
    return _protocol_4_verifier_ints(g_hat, k, Q, coeffs, gf, proof, round_i)


def _to_linear(L, y, n, gf, known=None):
    """pivot.affine_to_linear without evaluating L on n zeros one coefficient at a time: when every coefficient
    lives in gf, L(0, ..., 0) is gf(0) + L.constant (what the reference's sum() produces).  `known`: a one-element list
    that receives whether the coefficients were found to be gf elements (subtracting a constant keeps them so), which
    saves the caller a second pass over them."""
    in_field = FAST_INT_PATH and _all_in_field(L.coeffs, gf)
    if known is not None:
        known.append(bool(in_field))
    if in_field:
        constant = gf(0) + L.constant
        return L - constant, y - constant
    return pivot.affine_to_linear(L, y, n)


class _FormText:
    """repr-compatible stand-in for an affine form with gf coefficients: same text, built from the residues."""
    __slots__ = ("form", "q", "signed")

    def __init__(self, form, q, signed):
        self.form, self.q, self.signed = form, q, signed

    def __repr__(self):
        return f"{_field_text([c.value for c in self.form.coeffs], self.q, self.signed)}, {str(self.form.constant)}"

    def transcript_scalars(self):
        return len(self.form.coeffs), pivot.pack_scalars([c.value for c in self.form.coeffs], self.q), self.form.constant


def _first_prefix(t, A, generators, P, L, y, order, gf=None, L_text=None):
    """The O(N) part of the first pre-image (text or bytes of the generators and of L), hashed once: both challenges
    share it.  Only device calls and SHA-256 updates, both of which release the GIL -- protocol_5's device front end
    runs it on a worker thread while the main thread packs the witness."""
    if L_text is not None:
        L = L_text
    elif FAST_INT_PATH and gf is not None and gf.order == order and _all_in_field(L.coeffs, gf):
        L = _FormText(L, order, bool(gf.is_signed))
    input_list = [t, A.normalize(), generators, P.normalize(), L, y]
    if logger_cp_hin.isEnabledFor(logging.DEBUG):
        logger_cp_hin.debug(f"Before fiat_shamir_hash, input_list=\n{input_list}")
    if pivot.TRANSCRIPT == "binary":
        return "binary", pivot.binary_prefix(b"cp-first", input_list, order)
    return "reference", pivot.fiat_shamir_prefix(input_list)


def _first_finish(state, order):
    mode, prefix = state
    if mode == "binary":
        return pivot.binary_finish(prefix, [0], order), pivot.binary_finish(prefix, [1], order)
    # str(input_list + [b] + [tag]) for b = 0, 1 share everything but one character
    c0 = pivot.fiat_shamir_finish(prefix, [0, _TAG], order)
    c1 = pivot.fiat_shamir_finish(prefix, [1, _TAG], order)
    logger_cp_hout.debug(f"After hash, hash=\n{c0}, {c1}")
    return c0, c1


def _first_challenges(t, A, generators, P, L, y, order, gf=None, L_text=None):
    return _first_finish(_first_prefix(t, A, generators, P, L, y, order, gf, L_text), order)


def _g_hat(g, h, group):
    """Device copy of g + [h], owned by the caller (it is folded in place afterwards)."""
    dev = pivot.as_device_list(g, group)
    hd = DevicePointList(group, pivot._device_single(group, h))
    out = dev.clone(hd)
    out._owned = True
    return out


def _protocol_5_prover_dev(generators, g_hat, P, L, y, x, gamma, gf, r, rho):
    """protocol_5_prover with every length-n vector on the device: r_ext = r + [rho], x_ext = x + [gamma] and
    L_ext = L.coeffs + [0] are uploaded once; t = <L, r>, A = g_hat^{r_ext}, z_hat = r_ext + c0 x_ext (its last entry
    is phi = rho + c0 gamma) and L_tilde = c1 L_ext are device operations.  Same values, same transcript."""
    k = generators["k"]
    group = type(k)
    order = gf.order
    n = len(x)
    ctx = g_hat.dev.ctx
    proof = {}
    _mark("p5:start")
    if isinstance(r, list):
        zd = ctx.upload_scalars(r + [rho], order)
    else:  # packed draws (pivot.random_residues_packed)
        zd = ctx.upload_scalars(r.tobytes() + (rho % order).to_bytes(32, "little"), order)
    Ld = xd = None
    try:
        # the announcement A only needs r: it is issued first and computed while the host packs the coefficients of L
        logger_cp.debug("Calculate A.")
        g_fixed = generators["g"]
        if isinstance(g_fixed, DevicePointList) and getattr(g_fixed.dev, "precomputed", False) and len(g_fixed) >= n:
            # fixed generators with a table (DevicePointList.precompute): A through it, h as the extra term
            hd = pivot._device_single(group, generators["h"])
            ctx.msm_dev_ext(g_fixed.dev, g_fixed.off, n, zd, 0, hd, 0, [rho], slot=0)
        else:
            ctx.msm_dev(g_hat.dev, zd, slot=0, poff=g_hat.off, soff=0, n=n + 1)  # h**rho * prod g_i**r_i
        Ld = ctx.upload_scalars(_pack_field(L.coeffs, gf, order) + bytes(32), order)
        _mark("p5:uploads")
        t = gf(ctx.scalars_dot(Ld, 0, zd, 0, n)) + L.constant
        A = group._make(ctx.result(0))
        _mark("p5:t,A")
        proof["t"] = t
        proof["A"] = A
        # the first pre-image (text of the generators and of L: device calls + SHA-256, no GIL) is hashed on a worker
        # thread while this thread packs the witness, which is pure Python and touches no device state
        L_text = _DevForm(Ld, n, bool(gf.is_signed), L.constant)
        if n >= FIRST_HASH_THREAD_MIN:  # below that, starting a thread costs more than the overlap returns
            with _WORKER() as pool:
                pending = pool.submit(_first_prefix, t, A, generators, P, L, y, order, gf, L_text)
                x_raw = _pack_field(x, gf, order) + pivot.pack_scalars([pivot._int(gamma)], order)
                state = pending.result()
        else:
            state = _first_prefix(t, A, generators, P, L, y, order, gf, L_text)
            x_raw = _pack_field(x, gf, order) + pivot.pack_scalars([pivot._int(gamma)], order)
        _mark("p5:first hash + pack x")
        xd = ctx.upload_scalars(x_raw, order)
        c0, c1 = _first_finish(state, order)
        zd.axpy(c0, xd, _lib.AXPY_ADD_SCALED)
        logger_cp.debug("Calculate Q.")
        Q = group.lincomb([A, P, k], [1, c0, int(c1 * (c0 * y + t))])
        l_z = ctx.scalars_dot(Ld, 0, zd, 0, n)
        Ld.axpy(c1, None, _lib.AXPY_SCALE)
        assert l_z * c1 % order == ctx.scalars_dot(Ld, 0, zd, 0, n + 1)  # L(z) * c1 == L_tilde(z_hat)
        _mark("p5:z, Q, L_tilde")
    except BaseException:
        zd.free()
        if Ld is not None:
            Ld.free()
        raise
    finally:
        if xd is not None:
            xd.free()
    return _protocol_4_prover_dev(g_hat, k, Q, Ld, zd, gf, proof, 0)


def protocol_5_prover(generators, P, L, y, x, gamma, gf):
    """Compressed Sigma-protocol Pi_c, prover (reference :89-145)."""
    g, h, k = generators["g"], generators["h"], generators["k"]
    group = type(h)
    proof = {}
    n = len(x)
    coeffs_in_field = []
    L, y = _to_linear(L, y, n, gf, coeffs_in_field)
    assert bin(n + 1).count("1") == 1, \
        "This implementation requires n+1 to be power of 2 (else, use padding with zeros)."
    order = gf.order
    # the residue paths reduce modulo gf.order while the round loop reduces modulo the group order: same thing only when
    # the field IS the exponent field of the group (always so in the reference's drivers); otherwise the generic path
    fast = (FAST_INT_PATH and gf.order == k.order and coeffs_in_field[0] and _all_in_field(x, gf)
            and L.constant == 0)
    if fast and len(L.coeffs) == n:
        g_hat = _g_hat(g, h, group)
        if _use_device_scalars(g_hat, order, DEVICE_SCALAR_MIN_PROVER):
            # the announcement randomness only ever lives on the device here: drawn (same draws, same generator state
            # as the reference's n randrange calls, compressed_pivot.py:105-106) straight into packed bytes
            r = pivot.random_residues_packed(prng, order, n)
            if r is None:
                r = pivot.random_residues(prng, order, n)
            rho = prng.randrange(order)
            return _protocol_5_prover_dev(generators, g_hat, P, L, y, x, gamma, gf, r, rho)
        del g_hat
    r = pivot.random_residues(prng, order, n)
    rho = prng.randrange(order)
    logger_cp.debug("Calculate t.")
    t = gf(_dot([cf.value for cf in L.coeffs], r, order)) + L.constant if fast else L(r)
    logger_cp.debug("Calculate A.")
    A = pivot.vector_commitment(r, rho, g, h)
    proof["t"] = t
    proof["A"] = A
    c0, c1 = _first_challenges(t, A, generators, P, L, y, order, gf)
    phi = gf(c0 * gamma + rho)
    if fast:
        zi = [(c0 * x_i.value + r_i) % order for x_i, r_i in zip(x, r)] + [phi.value]
    else:
        z = [c0 * x_i + r_i for x_i, r_i in zip(x, r)]
        z_hat = z + [phi]
    g_hat = _g_hat(g, h, group)
    logger_cp.debug("Calculate Q.")
    Q = group.lincomb([A, P, k], [1, c0, int(c1 * (c0 * y + t))])
    if fast:
        # same values as the generic lines below; the appended coefficient 0 prints as "0" either way
        lc = [cf.value for cf in L.coeffs]
        coeffs = [v * c1 % order for v in lc] + [0]
        assert _dot(lc, zi[:-1], order) * c1 % order == _dot(coeffs, zi, order)
        return _protocol_4_prover_fast(g_hat, k, Q, coeffs, zi, gf, proof, 0)
    L_tilde = pivot.LinearForm(L.coeffs + [0]) * c1
    assert L(z) * c1 == L_tilde(z_hat)
    return protocol_4_prover(g_hat, k, Q, L_tilde, z_hat, gf, proof)


def protocol_4_verifier(g_hat, k, Q, L_tilde, gf, proof, round_i=0):
    """Non-interactive protocol 4, verifier (reference :148-202): the same folds plus one final 2-term commitment."""
    group = type(k)
    g_hat = _private_device_list(g_hat, group)
    order = k.order
    if (FAST_INT_PATH and isinstance(L_tilde, pivot.LinearForm) and gf.order == order
            and _all_in_field(L_tilde.coeffs, gf)):
        return _protocol_4_verifier_fast(g_hat, k, Q, [c.value for c in L_tilde.coeffs], gf, proof, round_i)
    while True:
        half = len(g_hat) // 2
        logger_cp.debug("Load from proof: A_i, B_i.")
        A = proof["A" + str(round_i)]
        B = proof["B" + str(round_i)]
        Q = _resolve(Q)
        c = _fold_challenge(A, B, g_hat, k, Q, L_tilde, order)
        g_hat = _fold_generators(g_hat, c)
        Q = _q_prime(group, A, Q, B, c)
        L_tilde = _fold_forms(L_tilde, c, half, gf)
        if len(g_hat) <= 2:
            z_prime = proof["z_prime"]
            Q_check = pivot.vector_commitment(z_prime, int(L_tilde(z_prime)), g_hat, k)
            Q = _resolve(Q)
            logger_cp.debug(f"Q_check= {Q_check}")
            logger_cp.debug(f"Q_prime= {Q}")
            return Q_check == Q
        round_i += 1


def protocol_5_verifier(generators, P, L, y, proof, gf):
    """Compressed Sigma-protocol Pi_c, verifier (reference :205-239)."""
    g, h, k = generators["g"], generators["h"], generators["k"]
    group = type(h)
    order = gf.order
    L, y = _to_linear(L, y, len(g), gf)
    logger_cp.debug("Load from proof: t, A.")
    t, A = proof["t"], proof["A"]
    g_hat = _g_hat(g, h, group)
    if (FAST_INT_PATH and gf.order == k.order and _all_in_field(L.coeffs, gf) and len(L.coeffs) + 1 == len(g_hat)
            and _use_device_scalars(g_hat, order)):
        Ld = g_hat.dev.ctx.upload_scalars(_pack_field(L.coeffs, gf, order) + bytes(32), order)
        try:
            c0, c1 = _first_challenges(t, A, generators, P, L, y, order, gf,
                                       L_text=_DevForm(Ld, len(L.coeffs), bool(gf.is_signed), L.constant))
            Q = group.lincomb([A, P, k], [1, c0, int(c1 * (c0 * y + t))])
            Ld.axpy(c1, None, _lib.AXPY_SCALE)
        except BaseException:
            Ld.free()
            raise
        return _protocol_4_verifier_dev(g_hat, k, Q, Ld, gf, proof, 0)
    c0, c1 = _first_challenges(t, A, generators, P, L, y, order, gf)
    Q = group.lincomb([A, P, k], [1, c0, int(c1 * (c0 * y + t))])
    if FAST_INT_PATH and gf.order == k.order and _all_in_field(L.coeffs, gf):
        coeffs = [cf.value * c1 % order for cf in L.coeffs] + [0]
        return _protocol_4_verifier_fast(g_hat, k, Q, coeffs, gf, proof, 0)
    L_tilde = pivot.LinearForm(L.coeffs + [0]) * c1
    return protocol_4_verifier(g_hat, k, Q, L_tilde, gf, proof)
