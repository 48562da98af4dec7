"""CPU: Groth16 keys made by oracle.zkey_new from a synthetic prepared ptau with known toxic waste (oracle/synth_setup.py)
give proofs that verify under the pairing check — on BN254 (same verifier that accepts the key derived from the reference's
ptau file) and on BLS12-381 (oracle/pairing_bls.py), where the reference ships no fixtures at all."""
import pytest

from oracle import oracle as orc
from oracle import synth_setup as ss
from oracle.plonk import wtns_bytes


@pytest.mark.parametrize("curve", [orc.BN254, orc.BLS12_381])
def test_groth16_structured_synthetic_key_verifies(curve):
    ci = orc.CURVES[curve]
    r1cs, wit = ss.chain_r1cs(curve, 100)
    ptau = ss.prepared_ptau(curve, 128, tau=0x1234567890ABCDEF, alpha=0xAAAA5555, beta=0xBBBB7777)
    zkey = orc.zkey_new(r1cs, ptau)
    zk = orc.read_zkey_header(*orc.read_binfile(zkey, "zkey", 2))
    assert (zk["nPublic"], zk["domainSize"]) == (1, 128)
    proof, public = orc.groth16_prove(zkey, wtns_bytes(wit, ci.r), ci.fr_to_mont(11), ci.fr_to_mont(13))
    vk = orc.zkey_vk(zkey)
    pub = [int(p) for p in public]
    assert orc.groth16_verify(vk, pub, proof)
    assert not orc.groth16_verify(vk, [pub[0] ^ 1], proof)
    assert not orc.groth16_verify(vk, [pub[0] + ci.r], proof)            # aliased public input (groth16_verify.js:41-46)
    bad = dict(proof)
    bad["pi_c"] = proof["pi_a"]
    assert not orc.groth16_verify(vk, pub, bad)
    # other (r, s): another valid proof
    p2, _ = orc.groth16_prove(zkey, wtns_bytes(wit, ci.r), ci.fr_to_mont(99), ci.fr_to_mont(7))
    assert p2 != proof and orc.groth16_verify(vk, pub, p2)
