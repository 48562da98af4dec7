"""GPU parity of the PLONK prover (sb_plonk_load / sb_plonk_prove through the C ABI) against oracle/plonk.py:
identical proof objects for identical blinders, on the reference's own fixture key and on synthetic structured keys
(which also verify), with and without MSM window tables, plus the reference's error texts."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BLINDERS = [0x3000 + 104729 * i for i in range(11)]


@pytest.fixture(scope="module")
def env():
    import snarkjs_b200
    from oracle import oracle as orc
    from oracle import plonk as oplonk
    curve = snarkjs_b200.getCurveFromName("bn128")
    ci = orc.CURVES[orc.BN254]
    yield {"sb": snarkjs_b200, "orc": orc, "op": oplonk, "curve": curve, "bl": b"".join(ci.fr_to_mont(b) for b in BLINDERS)}
    curve.terminate()


def test_plonk_reference_fixture(env, golden):
    g = golden("plonk_case.npz")
    zkey, wtns = bytes(g["zkey"]), bytes(g["wtns"])
    pk = env["sb"].plonk.ProvingKey(zkey, env["curve"])
    try:
        proof, public = env["sb"].plonk.prove(pk, wtns, env["bl"])
        want, wpub = env["op"].plonk_prove(zkey, wtns, BLINDERS)
        assert public == wpub == json.loads(bytes(g["public_json"]))
        assert proof == want
        assert env["op"].plonk_verify(json.loads(bytes(g["vk_json"])), public, proof)
        # a second proof on the same key with other blinders: different proof, still valid
        p2, _ = env["sb"].plonk.prove(pk, wtns)
        assert p2 != proof and env["op"].plonk_verify(json.loads(bytes(g["vk_json"])), public, p2)
    finally:
        pk.release()


@pytest.mark.parametrize("n_gates,structured", [(13, True), (120, True), (1000, True), (4090, True), (16000, False)])
def test_plonk_synthetic(env, n_gates, structured):
    """4090 gates -> domain 4096, 4102 PTau points: the MSMs run in table mode; 16000 -> domain 2^14 on unstructured points."""
    op = env["op"]
    gates, adds, n_vars, n_pub, wit = op.chain_gates(n_gates)
    zkey = op.plonk_setup_synth(gates, adds, n_vars, n_pub, tau=0x5EED5EED5EED + n_gates, structured=structured)
    wtns = op.wtns_bytes(wit)
    pk = env["sb"].plonk.ProvingKey(zkey, env["curve"])
    try:
        assert (pk.nVars, pk.nAdditions) == (n_vars, len(adds))
        proof, public = env["sb"].plonk.prove(pk, wtns, env["bl"])
        want, wpub = op.plonk_prove(zkey, wtns, BLINDERS)
        assert public == wpub
        assert proof == want
        if structured and n_gates <= 1000:
            assert op.plonk_verify(op.plonk_vk(zkey), public, proof)
        # same key, same inputs -> same bytes (no state leaks between proofs)
        again, _ = env["sb"].plonk.prove(pk, wtns, env["bl"])
        assert again == proof
    finally:
        pk.release()


@pytest.mark.parametrize("n_gates,n_pub,with_additions", [(29, 3, True), (60, 5, False)])
def test_plonk_shapes(env, n_gates, n_pub, with_additions):
    """Several public inputs (PI(X) sums several Lagrange polynomials) and a key without additions."""
    op = env["op"]
    gates, adds, n_vars, n_pub, wit = op.chain_gates(n_gates, n_pub=n_pub, with_additions=with_additions)
    zkey = op.plonk_setup_synth(gates, adds, n_vars, n_pub, tau=31337 + n_gates)
    wtns = op.wtns_bytes(wit)
    pk = env["sb"].plonk.ProvingKey(zkey, env["curve"])
    try:
        proof, public = env["sb"].plonk.prove(pk, wtns, env["bl"])
        assert (proof, public) == op.plonk_prove(zkey, wtns, BLINDERS)
        assert len(public) == n_pub and op.plonk_verify(op.plonk_vk(zkey), public, proof)
    finally:
        pk.release()


def test_plonk_reference_circuit2(env, golden, reference_plonk_key):
    """The reference's larger PLONK key (test/circuit2: domain 2048, 1001 additions, 4 public signals; rebuilt byte for byte
    from its r1cs) with the reference's own witness."""
    zkey, wtns = reference_plonk_key(golden("plonk_setup_cases.npz"), "c2048")
    pk = env["sb"].plonk.ProvingKey(zkey, env["curve"])
    try:
        proof, public = env["sb"].plonk.prove(pk, wtns, env["bl"])
        assert (proof, public) == env["op"].plonk_prove(zkey, wtns, BLINDERS)
        assert len(public) == 4 and env["op"].plonk_verify(env["op"].plonk_vk(zkey), public, proof)
    finally:
        pk.release()


def test_plonk_deep_addition_chain(env):
    """Every addition depends on the previous one: one k_pl_additions launch per addition (dependency levels)."""
    op = env["op"]
    gates, adds, n_vars, n_pub, wit = op.chain_gates(100, deep_additions=True)
    zkey = op.plonk_setup_synth(gates, adds, n_vars, n_pub, tau=424242)
    wtns = op.wtns_bytes(wit)
    pk = env["sb"].plonk.ProvingKey(zkey, env["curve"])
    try:
        assert env["sb"].plonk.prove(pk, wtns, env["bl"]) == op.plonk_prove(zkey, wtns, BLINDERS)
    finally:
        pk.release()


def test_plonk_bls12381(env):
    """BLS12-381: 12-limb base field (transcript, MSM), its own Fr roots; parity with the oracle, and the proof verifies."""
    sb, op, orc = env["sb"], env["op"], env["orc"]
    ci = orc.CURVES[orc.BLS12_381]
    gates, adds, n_vars, n_pub, wit = op.chain_gates(120, r=ci.r)
    zkey = op.plonk_setup_synth(gates, adds, n_vars, n_pub, tau=99991, curve=orc.BLS12_381)
    wtns = op.wtns_bytes(wit, ci.r)
    curve = sb.getCurveFromName("bls12381")
    try:
        pk = sb.plonk.ProvingKey(zkey, curve)
        proof, public = sb.plonk.prove(pk, wtns, b"".join(ci.fr_to_mont(b) for b in BLINDERS))
        pk.release()
        want, wpub = op.plonk_prove(zkey, wtns, BLINDERS)
        assert (proof, public) == (want, wpub)
        assert op.plonk_verify(op.plonk_vk(zkey), public, proof)
    finally:
        curve.terminate()


def test_plonk_key_from_file(env, golden, tmp_path):
    """sb_plonk_load_file: the key mapped from disk gives the same proof as the key loaded from bytes."""
    g = golden("plonk_case.npz")
    zkey, wtns = bytes(g["zkey"]), bytes(g["wtns"])
    path = tmp_path / "circuit.zkey"
    path.write_bytes(zkey)
    pk = env["sb"].plonk.ProvingKey.from_file(str(path), env["curve"])
    try:
        ref = env["op"].read_plonk_zkey(zkey)
        assert (pk.nVars, pk.nPublic, pk.domainSize, pk.nAdditions) == (ref["nVars"], ref["nPublic"], ref["domainSize"], ref["nAdditions"])
        proof, public = env["sb"].plonk.prove(pk, wtns, env["bl"])
        assert (proof, public) == env["op"].plonk_prove(zkey, wtns, BLINDERS)
    finally:
        pk.release()
    with pytest.raises(env["sb"].SbError, match="cannot open"):
        env["sb"].plonk.ProvingKey.from_file(str(tmp_path / "missing.zkey"), env["curve"])


def test_plonk_errors(env, golden):
    sb, op, orc = env["sb"], env["op"], env["orc"]
    gates, adds, n_vars, n_pub, wit = op.chain_gates(60)
    zkey = op.plonk_setup_synth(gates, adds, n_vars, n_pub, tau=777)
    pk = sb.plonk.ProvingKey(zkey, env["curve"])
    try:
        bad = list(wit)
        bad[5] = (bad[5] + 1) % orc.P_BN_R
        with pytest.raises(sb.SbError, match="Copy constraints does not match|Polynomial is not divisible"):
            sb.plonk.prove(pk, op.wtns_bytes(bad), env["bl"])
        with pytest.raises(sb.SbError, match=r"Invalid witness length. Circuit: \d+, witness: \d+, \d+"):
            sb.plonk.prove(pk, op.wtns_bytes(wit[:-1]), env["bl"])
        # the key still works after the errors
        proof, public = sb.plonk.prove(pk, op.wtns_bytes(wit), env["bl"])
        assert proof == op.plonk_prove(zkey, op.wtns_bytes(wit), BLINDERS)[0]
    finally:
        pk.release()
    g16 = bytes(golden("groth16_case.npz")["zkey"])
    with pytest.raises(sb.SbError, match="zkey file is not plonk"):
        sb.plonk.ProvingKey(g16, env["curve"])
