"""CPU checks of the fflonk restatement in oracle/fflonk.py (its pins are listed in that file's header)."""
import copy
import json

import pytest

from oracle import fflonk
from oracle import oracle as orc

BLINDERS = [0x4000 + 1299709 * i for i in range(9)]


@pytest.fixture(scope="module")
def ref_case(golden):
    return {k: bytes(v) for k, v in golden("fflonk_case.npz").items()}


def test_vk_from_reference_zkey_matches_reference_vk(ref_case):
    assert fflonk.fflonk_vk(ref_case["zkey"]) == json.loads(ref_case["vk_json"])


def test_prove_reference_zkey_verifies_with_reference_vk(ref_case):
    vk = json.loads(ref_case["vk_json"])
    proof, public = fflonk.fflonk_prove(ref_case["zkey"], ref_case["wtns"], BLINDERS)
    assert public == json.loads(ref_case["public_json"])
    assert list(proof["polynomials"]) == ["C1", "C2", "W1", "W2"]
    assert list(proof["evaluations"]) == list(fflonk.EVAL_NAMES) + ["inv"]
    assert fflonk.fflonk_verify(vk, public, proof)
    for key in ("C1", "C2", "W1", "W2"):
        bad = copy.deepcopy(proof)
        bad["polynomials"][key] = proof["polynomials"]["C1" if key != "C1" else "C2"]
        assert not fflonk.fflonk_verify(vk, public, bad), key
    for key in ("ql", "s3", "a", "zw", "t2w"):
        bad = copy.deepcopy(proof)
        bad["evaluations"][key] = str((int(proof["evaluations"][key]) + 1) % orc.P_BN_R)
        assert not fflonk.fflonk_verify(vk, public, bad), key
    assert not fflonk.fflonk_verify(vk, [str(int(public[0]) ^ 1)], proof)
    assert not fflonk.fflonk_verify(vk, [str(int(public[0]) + orc.P_BN_R)], proof)       # aliased public input
    # `inv` is the inverse of the product the on-chain verifier would otherwise have to invert (:1182-1245): not used
    # by the JS verifier, but it must be a field element and differ between proofs with different challenges
    p2, _ = fflonk.fflonk_prove(ref_case["zkey"], ref_case["wtns"], [b + 1 for b in BLINDERS])
    assert fflonk.fflonk_verify(vk, public, p2) and p2["evaluations"]["inv"] != proof["evaluations"]["inv"]


def test_wrong_witness_is_rejected(ref_case):
    _, w = orc.read_wtns(ref_case["wtns"])
    wit = [int.from_bytes(w[i:i + 32], "little") for i in range(0, len(w), 32)]
    wit[3] = (wit[3] + 1) % orc.P_BN_R
    from oracle.plonk import wtns_bytes
    with pytest.raises(ValueError):
        fflonk.fflonk_prove(ref_case["zkey"], wtns_bytes(wit), BLINDERS)


@pytest.mark.parametrize("n_gates,n_pub", [(13, 1), (120, 3)])
def test_synthetic_setup_prove_verify(n_gates, n_pub):
    from oracle import plonk
    gates, adds, n_vars, n_pub, wit = plonk.chain_gates(n_gates, n_pub=n_pub)
    zkey = fflonk.fflonk_setup_synth(gates, adds, n_vars, n_pub, tau=0xC0FFEE12345)
    proof, public = fflonk.fflonk_prove(zkey, plonk.wtns_bytes(wit), BLINDERS)
    vk = fflonk.fflonk_vk(zkey)
    assert len(public) == n_pub and fflonk.fflonk_verify(vk, public, proof)
    assert not fflonk.fflonk_verify(vk, [str(int(public[0]) ^ 1)] + public[1:], proof)


def test_fflonk_setup_reproduces_reference_zkey_byte_for_byte(golden, ref_case):
    """oracle.fflonk.fflonk_setup (r1cs -> gates and additions through r1cs_constraint_processor.js, selectors, sigmas with the
    two blinding rows, Lagrange, PTau, C0, header) gives exactly test/fflonk/circuit.zkey (593 092 bytes, 100 additions)."""
    g = golden("plonk_setup_cases.npz")
    ptau = orc.write_binfile("ptau", 1, [(1, bytes(g["ptau_header"])), (2, bytes(g["ff256_ptau2"])), (3, bytes(g["ptau3"])), (12, b"")])
    assert fflonk.fflonk_setup(bytes(g["ff256_r1cs"]), ptau) == ref_case["zkey"]
