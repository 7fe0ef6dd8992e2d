"""mpyc.secgroups look-alike (name only; mpc_ac20.py:12 imports it at module load)."""


def repeat_public_base_public_output(a, x):
    raise NotImplementedError("MPyC secure groups are not part of the test shim")
