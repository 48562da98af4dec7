"""CPU: the pure-Python parts of the host mirrors (snarkjs_b200/plonk.py, fflonk.py) — header readers and the
raw-bytes -> proof-object conversion — checked against the oracle without a GPU (the raw bytes come from the host-backend
flows of tests/host/)."""
import os
import sys
from types import SimpleNamespace

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import fflonk as off
from oracle import oracle as orc
from oracle import plonk as op
from snarkjs_b200 import fflonk as sb_fflonk
from snarkjs_b200 import plonk as sb_plonk
from snarkjs_b200.curve import SbError

import test_host_fflonk as HF
import test_host_plonk as HP

CURVE = SimpleNamespace(name="bn128", n8q=32, q=orc.P_BN_Q, r=orc.P_BN_R)


def _build(tmp_path_factory, src, name):
    import ctypes, os, subprocess
    so = str(tmp_path_factory.mktemp(name) / f"lib{name}.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(HP.ROOT, "tests", "host", src), "-ldl"])
    lib = ctypes.CDLL(so)
    for fn in ("hp_plonk_prove", "hp_fflonk_prove"):
        if hasattr(lib, fn):
            getattr(lib, fn).restype = ctypes.c_int
            getattr(lib, fn).argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_uint64,
                                         ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    return lib


def test_plonk_mirror_header_and_proof_object(tmp_path_factory, golden):
    g = golden("plonk_case.npz")
    zkey, wtns = bytes(g["zkey"]), bytes(g["wtns"])
    hdr = sb_plonk.read_zkey_header_plonk(zkey)
    ref = op.read_plonk_zkey(zkey)
    for k in ("nVars", "nPublic", "domainSize", "nAdditions", "nConstraints", "power", "q", "r"):
        assert hdr[k] == ref[k], k
    lib = _build(tmp_path_factory, "host_plonk.cpp", "hp")
    rc, err, raw = HP.host_prove(lib, zkey, wtns, HP.BLINDERS)
    assert rc == 0, err
    want, _ = op.plonk_prove(zkey, wtns, HP.BLINDERS)
    assert sb_plonk.proof_to_object(CURVE, raw) == want
    assert list(sb_plonk.proof_to_object(CURVE, raw)) == list(want)          # same key order as the reference's JSON
    with pytest.raises(SbError, match="zkey file is not plonk"):
        sb_plonk.read_zkey_header_plonk(bytes(golden("groth16_case.npz")["zkey"]))
    # the point at infinity is written [0, 1, 0] (G1.toObject)
    assert sb_plonk.proof_to_object(CURVE, bytes(len(raw)))["A"] == ["0", "1", "0"]


def test_fflonk_mirror_header_and_proof_object(tmp_path_factory, golden):
    g = golden("fflonk_case.npz")
    zkey, wtns = bytes(g["zkey"]), bytes(g["wtns"])
    hdr = sb_fflonk.read_zkey_header_fflonk(zkey)
    ref = off.read_fflonk_zkey(zkey)
    for k in ("nVars", "nPublic", "domainSize", "nAdditions", "nConstraints", "power", "q", "r"):
        assert hdr[k] == ref[k], k
    lib = _build(tmp_path_factory, "host_fflonk.cpp", "hf")
    rc, err, raw = HF.host_prove(lib, zkey, wtns, HF.BLINDERS)
    assert rc == 0, err
    want, _ = off.fflonk_prove(zkey, wtns, HF.BLINDERS)
    got = sb_fflonk.proof_to_object(CURVE, raw)
    assert got == want and list(got["evaluations"]) == list(want["evaluations"])
    with pytest.raises(SbError, match="zkey file is not fflonk"):
        sb_fflonk.read_zkey_header_fflonk(bytes(golden("plonk_case.npz")["zkey"]))


def _oracle_callbacks(ci):
    """The builders' field callables backed by the CPU oracle (what bench.py's reference arm uses)."""
    wn_of = lambda p: ci.fr_from_mont(orc.fr_root(ci.id, p))
    return (wn_of, lambda b, inv: orc.fr_fft(ci.id, b, inv), lambda b, f, i: orc.fr_batch_apply_key(ci.id, b, f, i),
            lambda grp, sd, k: orc.gen_points(ci.id, grp, sd, k), ci.g2_affine_bytes(ci.g2))


@pytest.mark.parametrize("curve_id,n_gates", [(orc.BN254, 26), (orc.BN254, 500), (orc.BLS12_381, 120)])
def test_synth_plonk_key_builder_equals_oracle_setup(curve_id, n_gates):
    """snarkjs_b200/synth.py builds the bench's PLONK keys with numpy + the library's NTT; with the oracle's NTT behind the
    same callables it must give the bytes of oracle.plonk.plonk_setup_synth (the restated plonk_setup.js) for the same gates."""
    from snarkjs_b200 import synth
    ci = orc.CURVES[curve_id]
    gates, adds, n_vars, n_pub, wit = op.chain_gates(n_gates, r=ci.r)
    want = op.plonk_setup_synth(gates, adds, n_vars, n_pub, tau=4242, structured=False, curve=curve_id)
    circ = synth.plonk_chain_circuit(n_gates, ci.r)
    assert circ["n_vars"] == n_vars and circ["witness"].tobytes() == b"".join(int(x).to_bytes(32, "little") for x in wit)
    got = synth.plonk_zkey_image(ci.q, ci.r, ci.n8q, circ, *_oracle_callbacks(ci), seed=4242)
    assert got == want


@pytest.mark.parametrize("n_gates", [26, 300])
def test_synth_fflonk_key_builder_equals_oracle_setup(n_gates):
    from snarkjs_b200 import synth
    ci = orc.CURVES[orc.BN254]
    gates, adds, n_vars, n_pub, wit = op.chain_gates(n_gates)
    want = off.fflonk_setup_synth(gates, adds, n_vars, n_pub, tau=4242, structured=False)
    circ = synth.plonk_chain_circuit(n_gates, ci.r)
    got = synth.fflonk_zkey_image(ci.q, ci.r, ci.n8q, circ, *_oracle_callbacks(ci), seed=4242)
    assert got == want
