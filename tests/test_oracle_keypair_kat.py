"""CPU: the reference's only hard-coded known-answer test (test/keypar_test.js:20-119, vectors in
tests/golden/keypair_kat.json) against the oracle: g2_sp is re-derived (blake2b -> ChaCha -> G2.fromRng, oracle/keypair.py)
and the three pairing equalities  e(g1_sx, g2_sp) == e(g1_s, g2_spx)  must hold under the oracle's own BN254 pairing —
the pairing every Groth16 / PLONK / fflonk verification in this suite relies on."""
import json
import os

import pytest

from oracle import oracle as O
from oracle import keypair as KP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = json.load(open(os.path.join(ROOT, "tests", "golden", "keypair_kat.json")))


def _g1(v): return (int(v[0], 16), int(v[1], 16))
def _g2(v): return ((int(v[0][0], 16), int(v[0][1], 16)), (int(v[1][0], 16), int(v[1][1], 16)))


def test_chacha_is_the_rfc_block_function():
    """The generator is ChaCha20's block function with a zero nonce and a 128-bit counter: first block of the all-zero key
    (RFC 7539 test vector 2.3.2 does not apply to a zero key; this is the well-known zero-key keystream, little-endian words)."""
    rng = KP.ChaCha([0] * 8)
    words = [rng.u32() for _ in range(4)]
    assert b"".join(w.to_bytes(4, "little") for w in words).hex() == "76b8e0ada0f13d90405d6ae55386bd28"


@pytest.mark.parametrize("case", KAT["cases"], ids=[c["name"] for c in KAT["cases"]])
def test_keypair_pairing_equalities(case):
    challenge = bytes.fromhex(KAT["challenge_hex"])
    s, sx, spx = _g1(case["g1_s"]), _g1(case["g1_sx"]), _g2(case["g2_spx"])
    q = O.P_BN_Q
    assert (s[1] ** 2 - s[0] ** 3 - 3) % q == 0 and (sx[1] ** 2 - sx[0] ** 3 - 3) % q == 0
    sp = KP.get_g2sp(case["personalization"], challenge, s, sx)
    # e(sx, sp) * e(-s, spx) == 1
    assert O.pairing_product_is_one([(sx, sp), ((s[0], (-s[1]) % q), spx)])
    # and the equality is not vacuous: a different personalization byte gives a point that breaks it
    other = KP.get_g2sp((case["personalization"] + 1) % 3, challenge, s, sx)
    assert other != sp
    assert not O.pairing_product_is_one([(sx, other), ((s[0], (-s[1]) % q), spx)])
