"""mpyc.mpctools look-alike: reduce() as a binary-tree reduction (call sites pivot.py:28, circuit_builder.py:299)."""
_no_value = object()


def reduce(f, x, initial=_no_value):
    x = list(x)
    if initial is not _no_value:
        x.insert(0, initial)
    if not x:
        raise TypeError("reduce() of empty sequence with no initial value")
    while len(x) > 1:
        odd = len(x) % 2
        x[odd:] = [f(x[i], x[i + 1]) for i in range(odd, len(x), 2)]
    return x[0]
