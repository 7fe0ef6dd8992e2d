"""CPU oracle for the verifiable_mpc MSM / fold hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``verifiable_mpc_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may use it, and there only as the checker / the CPU baseline being timed.

Parity status: **parity unpinned at the MPyC boundary**.  The reference's arithmetic for this path
lives in the third-party package MPyC (``mpyc.fingroups`` / ``mpyc.finfields``, pinned only as
``mpyc >= 0.8`` in /root/reference/setup.py:28, CI uses git HEAD, .travis.yml:11-13), which is
absent from /root/reference and from this image, and the reference's own tests hold no golden
vectors for group operations.  The oracle therefore restates the *published*
algorithms (RFC 8032 curve constants, EFD add-2008-bbjlp / dbl-2008-bbjlp projective formulas,
right-to-left binary double-and-add, binary-tree product) and is pinned by
  * RFC 8032 known answers (base point, order, small multiples: tests/test_oracle_kat.py),
  * two independent formula sets (affine vs projective) agreeing after normalisation,
  * algebraic identities with known discrete logs, and
  * the reference's own prover/verifier (imported unmodified from /root/reference on top of
    ``oracle/mpyc_shim``) accepting -- fixtures under tests/golden/ are generated that way.
Bit-exactness is always defined on the canonical affine encoding (x, y) reduced mod p.
"""
