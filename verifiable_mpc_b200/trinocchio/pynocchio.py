"""GPU-backed twin of the PROVER step of ``verifiable_mpc/trinocchio/pynocchio.py``: ``compute_proof`` (:228-273).

The reference builds eight lists ``[int(c[i]) * evalkey[key_i] ...]`` (one Python double-and-add per term, six over
BN256 G1 and one over G2 for the mid-wire indices, one over G1 for the quotient polynomial h) and sums each with
``apply_to_list(point_add, ...)`` (:82-91), then adds up to nine single scalar multiplications for the zero-knowledge
shifts.  Here every one of the eight proof elements is ONE device MSM (libvmsm.so, Pippenger over BN256 Jacobian
arithmetic) with the zero-knowledge terms riding along as extra terms of the same MSM.  Same signature, same proof
keys; key generation and the pairing-based ``verify`` stay the reference's (north_star: "pairing verify unchanged").
"""
from .. import _lib, hostpack
from ..engine import BN_N, pack_scalars

# (proof key, evalkey key template for mid index i, delta terms [(delta attribute, evalkey key)])
_MID_SUMS = (
    ("r_v*v_mid*g1", "r_v*v{i}*g1", (("v", "r_v*t*g1"),)),
    ("r_w*w_mid*g2", "r_w*w{i}*g2", (("w", "r_w*t*g2"),)),
    ("r_y*y_mid*g1", "r_y*y{i}*g1", (("y", "r_y*t*g1"),)),
    ("r_v*alpha_v*v_mid*g1", "r_v*alpha_v*v{i}*g1", (("v", "r_v*alpha_v*t*g1"),)),
    ("r_w*alpha_w*w_mid*g1", "r_w*alpha_w*w{i}*g1", (("w", "r_w*alpha_w*t*g1"),)),
    ("r_y*alpha_y*y_mid*g1", "r_y*alpha_y*y{i}*g1", (("y", "r_y*alpha_y*t*g1"),)),
    ("r_v*beta*v_mid+r_w*beta*w_mid+r_y*beta*y_mid*g1", "r_v*beta*v+r_w*beta*w+r_y*beta*y{i}_g1",
     (("v", "r_v*beta*t*g1"), ("w", "r_w*beta*t*g1"), ("y", "r_y*beta*t*g1"))),
)


def apply_to_list(op, inputs):
    """Binary-tree application (reference :82-91); kept for callers that import it from this module."""
    n = len(inputs)
    if n == 1:
        return inputs[0]
    return op(apply_to_list(op, inputs[: n // 2]), apply_to_list(op, inputs[n // 2:]))


def point_add(a, b):
    return a @ b


def _pack(values):
    """Witness / quotient coefficients (ints or elements of GF(n)) -> packed residues mod n; one C loop when the
    hostpack helper is built, ``int(v) % n`` per element otherwise (what the reference's ``int(c[i]) * P`` sees)."""
    raw = hostpack.pack_auto(values, BN_N)
    return raw if raw is not None else pack_scalars([int(v) for v in values], BN_N)


def _msm(points, scalars):
    """sum_i scalars[i] * points[i] on the device; points are group elements of one BN256 group."""
    group = type(points[0])
    ctx = group._ctx()
    dev = ctx.upload_points([p.affine() for p in points], curve=group.curve_id)
    try:
        return group._make(ctx.msm(dev, pack_scalars(scalars, BN_N)))
    finally:
        dev.free()


class PreparedEvalKey:
    """The base vectors of the eight sums uploaded ONCE (they only depend on the QAP and the evaluation key), so a
    prover that produces many proofs for the same circuit pays H2D for the witness scalars only."""

    def __init__(self, qap, evalkey, h_len):
        self.indices_mid = list(qap.indices_mid)
        self.h_len = h_len
        self.groups, self.bases = {}, {}
        for name, template, deltas in _MID_SUMS:
            pts = [evalkey[template.format(i=i)] for i in self.indices_mid] + [evalkey[k] for _, k in deltas]
            self._put(name, pts)
        self._put("h*g1", [evalkey["s^" + str(i) + "*g1"] for i in range(h_len)])

    @classmethod
    def from_device_bases(cls, indices_mid, h_len, groups, bases):
        """Key whose base vectors are already on the device (``bases[name]``: DevicePoints holding the mid-wire entries
        followed by the zero-knowledge entries of that sum, in ``_MID_SUMS`` order; ``"h*g1"``: the s-powers), e.g.
        produced by ``generate_evalkey``-style fixed-base batches without a round trip through host objects."""
        self = cls.__new__(cls)
        self.indices_mid, self.h_len, self.groups, self.bases = list(indices_mid), h_len, dict(groups), dict(bases)
        return self

    def precompute(self, window_bits=0):
        """Key tables (``vmsm_points_precompute``: 2^(13 w) * P for every key entry, 20 levels): the evaluation key is
        fixed per circuit, so the ~250 doublings of each of the eight sums are paid once per key instead of once per
        proof.  Costs 20 x the key's size in HBM (2^14 mid wires: 7 x 21 MB + 42 MB for the G2 sum)."""
        for dev in self.bases.values():
            dev.precompute(window_bits)
        return self

    def _put(self, name, pts):
        group = type(pts[0])
        self.groups[name] = group
        self.bases[name] = group._ctx().upload_points([p.affine() for p in pts], curve=group.curve_id)


def _evalkey_exponents(td, qap, order):
    """(key, exponent) pairs of every evaluation-key entry, split by group: each entry of the reference's
    ``generate_evalkey`` (:101-167) is ``int(alpha * poly(s)) * (r * g)`` = a multiple of the group generator, so the
    whole key is two fixed-base batches.  The polynomial evaluations at the trapdoor point stay the reference's host
    algebra (one ``poly.eval(s)`` per wire)."""
    s = td.s
    mid = list(qap.indices_mid)

    def at_s(poly):
        return int(poly.eval(s)) % order

    v_s = {i: at_s(qap.v[i]) for i in mid}
    w_s = {i: at_s(qap.w[i]) for i in mid}
    y_s = {i: at_s(qap.y[i]) for i in mid}
    t_s = at_s(qap.t)
    r_v, r_w, r_y = td.r_v % order, td.r_w % order, td.r_y % order
    g1, g2 = [], []
    g1 += [(f"r_v*v{i}*g1", v_s[i] * r_v) for i in mid]
    g2 += [(f"r_w*w{i}*g2", w_s[i] * r_w) for i in mid]
    g1 += [(f"r_y*y{i}*g1", y_s[i] * r_y) for i in mid]
    g1 += [(f"r_v*alpha_v*v{i}*g1", td.alpha_v * v_s[i] % order * r_v) for i in mid]
    g1 += [(f"r_w*alpha_w*w{i}*g1", td.alpha_w * w_s[i] % order * r_w) for i in mid]
    g1 += [(f"r_y*alpha_y*y{i}*g1", td.alpha_y * y_s[i] % order * r_y) for i in mid]
    g1 += [(f"s^{i}*g1", pow(s, i, order)) for i in range(0, qap.d + 1)]
    g1 += [(f"r_v*beta*v+r_w*beta*w+r_y*beta*y{i}_g1",
            td.beta * v_s[i] % order * r_v + td.beta * w_s[i] % order * r_w + td.beta * y_s[i] % order * r_y) for i in mid]
    g1 += [("r_v*t*g1", t_s * r_v), ("r_y*t*g1", t_s * r_y),
           ("r_v*alpha_v*t*g1", td.alpha_v * t_s % order * r_v), ("r_w*alpha_w*t*g1", td.alpha_w * t_s % order * r_w),
           ("r_y*alpha_y*t*g1", td.alpha_y * t_s % order * r_y), ("r_v*beta*t*g1", td.beta * t_s % order * r_v),
           ("r_w*beta*t*g1", td.beta * t_s % order * r_w), ("r_y*beta*t*g1", td.beta * t_s % order * r_y),
           ("t*g1", t_s)]
    g2 += [("r_w*t*g2", t_s * r_w)]
    return [(k, e % order) for k, e in g1], [(k, e % order) for k, e in g2]


# the reference's dict order (:155-165): v, w, y, alpha_v, alpha_w, alpha_y, s-powers, beta, then the ZK elements
_ZK_KEYS = ("r_v*t*g1", "r_w*t*g2", "r_y*t*g1", "r_v*alpha_v*t*g1", "r_w*alpha_w*t*g1", "r_y*alpha_y*t*g1",
            "r_v*beta*t*g1", "r_w*beta*t*g1", "r_y*beta*t*g1", "t*g1")


def generate_evalkey(td, qap, gen):
    """Public evaluation key (reference :101-167) with the ~8|mid| + d + 11 scalar multiplications done as two
    fixed-base batches on the device (``vmsm_points_fixed_base`` on BN256 G1 / G2).  ``gen.g1`` / ``gen.g2`` must be the
    groups' standard generators (what ``Generators(td, group.generator, twist.generator)`` holds, :61-69)."""
    group1, group2 = type(gen.g1), type(gen.g2)
    assert gen.g1 == group1.generator and gen.g2 == group2.generator, "fixed-base tables are built for the standard generators"
    order = group1.order
    e1, e2 = _evalkey_exponents(td, qap, order)
    out = {}
    for group, entries in ((group1, e1), (group2, e2)):
        dev = group._ctx().fixed_base(scalars=[e for _, e in entries], curve=group.curve_id)
        try:
            for (key, _), pt in zip(entries, dev.tolist()):
                out[key] = group._make(pt)
        finally:
            dev.free()
    mid = list(qap.indices_mid)
    order_keys = ([f"r_v*v{i}*g1" for i in mid] + [f"r_w*w{i}*g2" for i in mid] + [f"r_y*y{i}*g1" for i in mid]
                  + [f"r_v*alpha_v*v{i}*g1" for i in mid] + [f"r_w*alpha_w*w{i}*g1" for i in mid]
                  + [f"r_y*alpha_y*y{i}*g1" for i in mid] + [f"s^{i}*g1" for i in range(0, qap.d + 1)]
                  + [f"r_v*beta*v+r_w*beta*w+r_y*beta*y{i}_g1" for i in mid] + list(_ZK_KEYS))
    return {k: out[k] for k in order_keys}


RESIDENT_WITNESS = True  # prepared keys: one upload of the witness shared by the seven mid-wire sums

TRACE = None  # tools/bench_bn256.py --trace sets a list: (label, perf_counter()) marks inside compute_proof


def _mark(label):
    if TRACE is not None:
        import time

        TRACE.append((label, time.perf_counter()))


def compute_proof(qap, c, h, evalkey, deltas=None):
    """Pinocchio proof elements for witness ``c`` and quotient polynomial ``h`` (reference :228-273).

    ``evalkey``: the reference's dict of group elements, or a ``PreparedEvalKey`` (bases resident on the device).
    All eight MSMs are issued asynchronously (the witness scalars are packed once and shared by the seven mid-wire
    sums) and fetched afterwards, so their latency-bound tails overlap on the device.
    """
    import ctypes

    prepared = evalkey if isinstance(evalkey, PreparedEvalKey) else None
    mid = prepared.indices_mid if prepared else list(qap.indices_mid)
    if prepared:
        assert len(h) <= prepared.h_len, "Not enough generators."
    _mark("start")
    c_mid_raw = _pack([c[i] for i in mid])
    _mark("packed witness")

    jobs, keep, proof = [], [], {}  # jobs: (name, group, device points, owned?)
    witness = {}  # context id -> the packed witness as a device-resident scalar vector (prepared keys)

    def issue_resident(name, group, dev, extra_scalars):
        """Mid-wire sum over a prepared key: the witness is uploaded ONCE and read in place by all seven sums; a sum's
        zero-knowledge terms continue the key vector and bring their scalars along (vmsm_msm_dev_ext).  No per-sum
        staging of 32 bytes x |mid| from pageable host memory, which made every issue wait for the sum two before it."""
        ctx = group._ctx()
        sc = witness.get(id(ctx))
        if sc is None:
            sc = witness[id(ctx)] = ctx.upload_scalars(c_mid_raw, BN_N)
        jobs.append((name, group, dev, False))
        slot = len(jobs) - 1
        if extra_scalars:
            ctx.msm_dev_ext(dev, 0, len(mid), sc, 0, dev, len(mid), extra_scalars, slot=slot)
        else:
            ctx.msm_dev(dev, sc, slot=slot, poff=0, soff=0, n=len(mid))

    def issue(name, group, dev, owned, raw):
        buf = ctypes.create_string_buffer(raw, len(raw)) if raw else ctypes.create_string_buffer(1)
        keep.append(buf)  # msm_async reads the host buffer until its result is fetched
        jobs.append((name, group, dev, owned))
        group._ctx().msm_async(dev, ctypes.cast(buf, ctypes.c_void_p), 0, len(raw) // 32, len(jobs) - 1)

    try:
        # the G2 sum first: its tail is the longest and then overlaps the other seven
        for name, template, delta_terms in sorted(_MID_SUMS, key=lambda t: not t[0].endswith("g2")):
            extra = [int(getattr(deltas, attr)) for attr, _ in delta_terms] if deltas is not None else []
            raw = c_mid_raw + pack_scalars(extra, BN_N) if extra else c_mid_raw
            if prepared and RESIDENT_WITNESS and mid:
                issue_resident(name, prepared.groups[name], prepared.bases[name], extra)
            elif prepared:
                issue(name, prepared.groups[name], prepared.bases[name], False, raw)
            else:
                pts = [evalkey[template.format(i=i)] for i in mid]
                if deltas is not None:
                    pts += [evalkey[k] for _, k in delta_terms]
                group = type(pts[0])
                issue(name, group, group._ctx().upload_points([p.affine() for p in pts], curve=group.curve_id), True, raw)
        _mark("issued mid sums")
        # the coefficients of h are packed while the device works on the seven sums already issued
        raw_h = _pack(h.coeffs[:len(h)] if isinstance(h.coeffs, list) else [h.coeffs[i] for i in range(0, len(h))])
        _mark("packed h")
        if prepared:
            issue("h*g1", prepared.groups["h*g1"], prepared.bases["h*g1"], False, raw_h)
        else:
            pts = [evalkey["s^" + str(i) + "*g1"] for i in range(0, len(h))]
            group = type(pts[0])
            issue("h*g1", group, group._ctx().upload_points([p.affine() for p in pts], curve=group.curve_id), True, raw_h)
        _mark("issued h sum")
        for slot, (name, group, dev, owned) in enumerate(jobs):
            proof[name] = group._make(group._ctx().result(slot, curve=group.curve_id))
            _mark("result " + name)
    finally:
        for name, group, dev, owned in jobs:
            if owned:
                dev.free()
        for sc in witness.values():
            sc.free()
    # same key order as the reference's dict
    return {name: proof[name] for name in [t[0] for t in _MID_SUMS] + ["h*g1"]}
