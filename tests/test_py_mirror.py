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
