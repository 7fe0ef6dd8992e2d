"""mpyc.thresha look-alike: Lagrange recombination vector (demos/demo_zkp_trinocchio.py:18)."""


def _recombination_vector(field, xs, x_r):
    xs = [int(x) for x in xs]
    vec = []
    for i, x_i in enumerate(xs):
        num, den = 1, 1
        for j, x_j in enumerate(xs):
            if i != j:
                num *= x_r - x_j
                den *= x_i - x_j
        vec.append(field(num) / field(den))
    return vec
