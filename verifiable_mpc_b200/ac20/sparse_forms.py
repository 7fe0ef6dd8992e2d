"""Sparse construction of the AC20 circuit forms: drop-in twins of
``circuit_builder.calculate_fg_form`` / ``calculate_h_form`` / ``calculate_circuit_forms`` / ``lagrange``
(verifiable_mpc/ac20/circuit_builder.py:417-544) and of ``Circuit.multiplication_triples`` (:133-151).

Why: the reference builds, for EVERY multiplication gate, a dense affine form over all n + m wires
(``construct_affine_form``, one list of n + m zeros per recursion step, child gates found by a linear scan of
``circuit.gates``) and then sums m dense forms of length n + 3 + 2m; ``lagrange`` goes through
``recombine._recombination_vectors``, an O(m^2) double loop (recombine.py:15-32).  At m = 512 that is ~0.1 s and half a
million list slots per wire, at m = 2^15 minutes and 2^31 slots, so no real circuit of the sizes the
device handles can reach ``protocol_5_prover``.  The forms themselves are sparse -- a wire of a multiplication gate
depends on a handful of inputs / earlier products -- so here

* a wire form is a dict {position: coefficient} + constant, memoised per gate (shared sub-expressions are walked once,
  children are found through a name index);
* ``calculate_fg_form`` accumulates  sum_j lambda_{j+1} * form_j  straight into ONE dense coefficient list;
* the Lagrange vector over the nodes 0..M-1 at the challenge is computed in O(M) from factorials and prefix / suffix
  products with a single modular inversion.

Results are the reference's objects: the same ``AffineForm`` / ``LinearForm`` coefficient VALUES AND TYPES (the
reference's integer arithmetic is unreduced -- coefficients are exact Python ints, and their decimal text enters the
Fiat-Shamir pre-images of circuit_sat_cb.py:149-162 -- so the sums here are the same exact integers), hence the same
transcript and the same proof.  Host-side integer bookkeeping only: nothing here touches the device.

The circuit argument is the reference's ``Circuit`` (duck-typed: ``gates``, ``mul_gates()``, ``input_ct``, ``mul_ct``,
``output_gates``; gates with ``op`` (an enum whose ``.name`` is add / mul / scalar_mul), ``inputs``, ``output.name``,
``mul_index``; variables with ``name``, ``input_index``).  INTEGRATION.md shows the rebinding.
"""
# The classes of the forms returned.  Bound into the reference's drivers (which add these forms to their own), rebind
# them to the reference's:  sparse_forms.AffineForm, sparse_forms.LinearForm = pivot.AffineForm, pivot.LinearForm
from .forms import AffineForm, LinearForm


def lagrange_vector(modulus, count, c):
    """Lagrange coefficients at ``c`` for the nodes 0, 1, ..., count-1, as residues in [0, modulus): the row
    ``_recombination_vectors(field, range(count), (c,))[0]`` of recombine.py:5-32, in O(count) instead of O(count^2).

    lambda_i = prod_{j != i} (c - j) / prod_{j != i} (i - j);  prod_{j != i} (i - j) = i! * (count-1-i)! * (-1)^(count-1-i).
    """
    q = modulus
    c %= q
    if count == 0:
        return []
    diffs = [(c - j) % q for j in range(count)]
    if 0 in diffs:  # c is one of the nodes: the reference divides by zero here as well
        raise ZeroDivisionError("challenge coincides with an interpolation node")
    # prefix[i] = prod_{j < i} (c - j), suffix[i] = prod_{j > i} (c - j)
    prefix = [1] * count
    for i in range(1, count):
        prefix[i] = prefix[i - 1] * diffs[i - 1] % q
    suffix = [1] * count
    for i in range(count - 2, -1, -1):
        suffix[i] = suffix[i + 1] * diffs[i + 1] % q
    fact = [1] * count
    for i in range(1, count):
        fact[i] = fact[i - 1] * i % q
    # denominators d_i = i! (count-1-i)! (-1)^(count-1-i); invert all of them with one pow
    dens = [fact[i] * fact[count - 1 - i] % q for i in range(count)]
    run = [1] * (count + 1)
    for i in range(count):
        run[i + 1] = run[i] * dens[i] % q
    inv = pow(run[count], -1, q)
    out = [0] * count
    for i in range(count - 1, -1, -1):
        d_inv = inv * run[i] % q
        inv = inv * dens[i] % q
        v = prefix[i] * suffix[i] % q * d_inv % q
        out[i] = (q - v) % q if (count - 1 - i) & 1 else v
    return out


def lagrange(gf, lagr_range, c):
    """Twin of ``circuit_builder.lagrange`` (:541-542) for node ranges 0..M-1 (the only ones the protocol uses)."""
    nodes = list(lagr_range)
    assert nodes == list(range(len(nodes))), "nodes must be 0, 1, ..., M-1"
    return lagrange_vector(gf.modulus, len(nodes), int(c))


class _Sparse:
    """coeffs: {position in the (input_ct + mul_ct)-vector: coefficient}, plus the constant of an affine form."""
    __slots__ = ("coeffs", "constant")

    def __init__(self, coeffs=None, constant=0):
        self.coeffs = coeffs if coeffs is not None else {}
        self.constant = constant

    def add(self, other):
        out = dict(self.coeffs)
        for k, v in other.coeffs.items():
            out[k] = out[k] + v if k in out else 0 + v
        return _Sparse(out, self.constant + other.constant)

    def scaled(self, s):
        return _Sparse({k: v * s for k, v in self.coeffs.items()}, self.constant * s)


class SparseCircuitForms:
    """Wire forms of one circuit, built once (construct_affine_form, circuit_builder.py:417-498, on sparse forms)."""

    def __init__(self, circuit):
        self.circuit = circuit
        self.n = circuit.input_ct
        self.m = circuit.mul_ct
        self._by_output = {g.output.name: g for g in circuit.gates}
        self._memo = {}

    @staticmethod
    def _is_var(x):
        return hasattr(x, "input_index") and hasattr(x, "name")

    def _wire(self, gate, wire):
        """The form of one input wire of ``gate`` (``construct_for_wire``)."""
        inp = gate.inputs[wire]
        if not self._is_var(inp):
            return _Sparse({}, 0 + inp)
        if inp.input_index is not None:
            return _Sparse({inp.input_index: 1}, 0)
        child = self._by_output[inp.name]
        kind = child.op.name
        if kind == "mul":
            return _Sparse({self.n + child.mul_index: 1}, 0)
        if kind in ("add", "scalar_mul"):
            return self._gate(child)
        raise ValueError(kind)

    def _gate(self, gate):
        """The form of the OUTPUT of an add / scalar_mul gate (``construct_affine_form(gate, circuit, None)``)."""
        key = id(gate)
        hit = self._memo.get(key)
        if hit is not None:
            return hit
        kind = gate.op.name
        if kind == "add":
            out = self._wire(gate, 0).add(self._wire(gate, 1))
        elif kind == "scalar_mul":
            if self._is_var(gate.inputs[0]):
                out = self._wire(gate, 0).scaled(gate.inputs[1])
            elif self._is_var(gate.inputs[1]):
                out = self._wire(gate, 1).scaled(gate.inputs[0])
            else:
                out = _Sparse({}, gate.inputs[0] * gate.inputs[1])
        elif kind == "mul":
            out = _Sparse({self.n + gate.mul_index: 1}, 0)
        else:
            raise ValueError(kind)
        self._memo[key] = out
        return out

    def wire_form(self, gate, wire):
        return self._wire(gate, wire) if wire is not None else self._gate(gate)

    # ---- positions: (inputs, products) -> the AC20 z-vector (x, f(0), g(0), h(0), h(1..2m)), convert_to_ac20 :501-514
    def _pos(self, k):
        return k if k < self.n else k + 3

    def dense(self, sparse, ac20=True):
        """AffineForm with the reference's dense coefficient list (ints 0 where nothing was added)."""
        length = self.n + 3 + 2 * self.m if ac20 else self.n + self.m
        coeffs = [0] * length
        for k, v in sparse.coeffs.items():
            coeffs[self._pos(k) if ac20 else k] = v
        return AffineForm(coeffs, sparse.constant)

    def multiplication_triples(self, inputs):
        """``Circuit.multiplication_triples`` (:133-151): left / right / output wire values of every mul gate, each wire
        evaluated through its sparse form (gates are in topological order, as the reference assumes)."""
        gamma = [0] * self.m
        alpha, beta = [0] * self.m, [0] * self.m

        def ev(sf):
            acc = 0
            for k, v in sf.coeffs.items():
                acc = acc + v * (inputs[k] if k < self.n else gamma[k - self.n])
            return acc + sf.constant

        for i, g in enumerate(self.circuit.mul_gates()):
            alpha[i] = ev(self._wire(g, 0))
            beta[i] = ev(self._wire(g, 1))
            gamma[i] = alpha[i] * beta[i]
        return alpha, beta, gamma


def calculate_fg_form(circuit, wire, challenge, gf, forms=None):
    """Twin of ``circuit_builder.calculate_fg_form`` (:517-530): the form of f(c) (wire 0) / g(c) (wire 1) in the
    coordinates of the z-vector.  ``forms``: a ``SparseCircuitForms`` to reuse between the two wires and calls."""
    sf = forms or SparseCircuitForms(circuit)
    n, m = sf.n, sf.m
    lam = lagrange(gf, range(m + 1), challenge)
    coeffs = [0] * (n + 3 + 2 * m)
    constant = 0
    coeffs[n + wire] = 1 * lam[0]
    for j, gate in enumerate(circuit.mul_gates()):
        l_j = lam[j + 1]
        form = sf.wire_form(gate, wire)
        for k, v in form.coeffs.items():
            p = sf._pos(k)
            coeffs[p] = coeffs[p] + v * l_j
        constant = constant + form.constant * l_j
    return AffineForm(coeffs, constant)


def calculate_h_form(circuit, challenge, gf):
    """Twin of ``circuit_builder.calculate_h_form`` (:533-537)."""
    lam = lagrange(gf, range(2 * circuit.mul_ct + 1), challenge)
    return LinearForm([0] * circuit.input_ct + [0] * 2 + lam)


def calculate_circuit_forms(circuit, forms=None):
    """Twin of ``circuit_builder.calculate_circuit_forms`` (:540-545): one form per output gate, over (inputs,
    products); the caller applies ``convert_to_ac20`` as the reference does (circuit_sat_cb.py:141-142)."""
    sf = forms or SparseCircuitForms(circuit)
    return [sf.dense(sf.wire_form(circuit.gates[ix], None), ac20=False) for ix in circuit.output_gates]
