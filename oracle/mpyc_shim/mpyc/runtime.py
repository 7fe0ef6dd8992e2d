"""mpyc.runtime look-alike: `logging` re-export and an `mpc` object that only exists (the asyncio MPC runtime is out
of scope; any use raises)."""
import logging  # noqa: F401  (re-exported: `from mpyc.runtime import logging`, pivot.py:16)


class _NoRuntime:
    def __getattr__(self, name):
        raise NotImplementedError(f"mpyc.runtime.mpc.{name}: the MPyC runtime is not part of the test shim")


mpc = _NoRuntime()
