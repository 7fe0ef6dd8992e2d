"""Pins the oracle: RFC 8032 known answers, curve/order checks, and agreement of its two formula sets.
(The reference's own tests hold no group-level golden vectors -- SURVEY.md F7 -- so standards are the anchor.)"""
import hashlib
import json
import os

from oracle import ed25519 as E
from oracle import prng

# RFC 8032 section 7.1, Ed25519 test vectors 1-3 (secret seed -> public key)
RFC8032 = [
    ("9d61b19deffd5a60ba844af492ec2cc44449c5697b326919703bac031cae7f60",
     "d75a980182b10ab7d54bfed3c964073a0ee172f3daa62325af021a68f707511a"),
    ("4ccd089b28ff96da9db6c346ec114e0f5b8a319f35aba624da8cf6ed4fb8a6fb",
     "3d4017c3e843895a92b70aa74d1b7ebc9c982ccf2ec4968cc0cd55f12af4660c"),
    ("c5aa8df43f9f837bedb7442f31dcb7b166d38535076f094b85ce3a2e0b4458f7",
     "fc51cd8e6218a1a38da47ed00230f0580816ed13ba3303ac5deb911548908025"),
]


def encode_rfc8032(pt):
    x, y = pt
    return (y | ((x & 1) << 255)).to_bytes(32, "little")


def test_constants():
    assert E.P == 2**255 - 19
    assert E.on_curve(E.B) and E.on_curve(E.IDENTITY)
    assert E.BY == 4 * pow(5, -1, E.P) % E.P  # RFC 8032: y = 4/5
    assert E.scalar_mul(E.B, E.L) == E.IDENTITY
    assert E.scalar_mul(E.B, E.L - 1) == E.affine_neg(E.B)


def test_rfc8032_public_keys():
    for seed_hex, pub_hex in RFC8032:
        h = hashlib.sha512(bytes.fromhex(seed_hex)).digest()
        a = int.from_bytes(h[:32], "little")
        a &= (1 << 254) - 8
        a |= 1 << 254
        assert encode_rfc8032(E.scalar_mul(E.B, a)).hex() == pub_hex
        assert encode_rfc8032(E.affine_repeat(E.B, a)).hex() == pub_hex


def test_formula_sets_agree():
    pts = [E.scalar_mul(E.B, prng.scalar(3, i)) for i in range(6)]
    for i, p in enumerate(pts):
        assert E.on_curve(p)
        q = pts[(i + 1) % 6]
        assert E.normalize(E.proj_add(E.to_projective(p), E.to_projective(q))) == E.affine_add(p, q)
        assert E.normalize(E.proj_dbl(E.to_projective(p))) == E.affine_add(p, p)
        k = prng.scalar(4, i)
        assert E.scalar_mul(p, k) == E.affine_repeat(p, k)
        assert E.scalar_mul(p, -k) == E.affine_neg(E.scalar_mul(p, k))
    assert E.scalar_mul(pts[0], 0) == E.IDENTITY


def test_msm_naive_vs_known_dlog():
    dl = [prng.scalar(5, i) for i in range(9)]
    pts = [E.scalar_mul(E.B, r) for r in dl]
    sc = [prng.scalar(6, i) for i in range(9)]
    assert E.msm_naive(sc, pts) == E.msm_known_dlog(sc, dl)
    # negative and unreduced scalars, as the reference passes them (pivot.py:119-128, compressed_pivot.py:66)
    sc2 = [-3, E.L + 5, sc[0] ** 2, 0, 1, -1, 2**300, 7, 8]
    assert E.msm_naive(sc2, pts) == E.msm_known_dlog(sc2, dl)
    assert E.vector_commitment(sc[:8], 12345, pts[:8], E.B) == E.msm_known_dlog(sc[:8] + [12345], dl[:8] + [1])


def test_fold_known_dlog():
    dl = [prng.scalar(8, i) for i in range(8)]
    pts = [E.scalar_mul(E.B, r) for r in dl]
    c = prng.scalar(9, 0)
    assert E.fold(pts, c) == [E.scalar_mul(E.B, (c * dl[i] + dl[4 + i]) % E.L) for i in range(4)]


def test_prng_spec():
    assert prng.scalar(0x5EED, 0) == 0xD29B6C7D22528D5D5BCA4696B343B340537A3778C7E79EB1DF9A82A6FAD5C7
    assert all(0 <= prng.scalar(1, i) < E.L for i in range(200))


def test_golden_fixture_matches_oracle():
    path = os.path.join(os.path.dirname(__file__), "golden", "ed25519_msm_fold.json")
    g = json.load(open(path))
    pts = [tuple(int(v, 16) for v in p) for p in g["points"]]
    for case in g["msm"]:
        sc = [int(s, 16) * (-1 if neg else 1) for s, neg in case["scalars"]]
        assert list(E.msm_naive(sc, pts)) == [int(v, 16) for v in case["expect"]]
    f = g["fold"]
    got = E.fold(pts[: f["n"]], int(f["c"], 16))
    assert [[hex(x), hex(y)] for x, y in got] == f["expect"]
