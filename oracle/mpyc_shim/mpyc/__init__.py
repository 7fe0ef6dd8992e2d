"""MPyC look-alike (TEST INFRASTRUCTURE ONLY) exposing exactly the names the reference imports,
so that /root/reference/verifiable_mpc/**.py can be imported UNMODIFIED in a container where the real MPyC
(third-party, `mpyc >= 0.8`, setup.py:28) is not installable.  Arithmetic is the pure-Python oracle.

Conventions that cannot be checked against real MPyC here (source absent) are marked UNVERIFIED in the modules;
most importantly `repr()` of an elliptic-curve point is defined as the canonical affine triple "[x, y, 1]" with
unsigned decimal coordinates, which makes Fiat-Shamir transcripts independent of the projective representative.
"""
__version__ = "0.0-shim"
