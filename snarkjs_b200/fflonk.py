"""fflonk.prove on the B200 — host-side mirror of src/fflonk_prove.js:51-267.

ProvingKey(zkey) puts the fflonk key in HBM once (sb_fflonk_load); prove() runs the five rounds on the device
(sb_fflonk_prove) and returns the {polynomials: {C1, C2, W1, W2}, evaluations: {ql .. t2w, inv}} object and the
publicSignals the reference writes (src/proof.js:62-82, fflonk_prove.js:240-262).  The blinders b_1..b_9 may be injected
as Montgomery field elements; otherwise they are drawn like Fr.random().  BN254 only, like the reference."""
from __future__ import annotations

import ctypes
import struct

import numpy as np

from .curve import Curve, SbError, _arr, _ptr, getCurveFromQ
from .groth16 import _from_mont, random_fr, read_binfile, read_wtns_header

POINTS = ("C1", "C2", "W1", "W2")
EVALS = ("ql", "qr", "qm", "qo", "qc", "s1", "s2", "s3", "a", "b", "c", "z", "zw", "t1w", "t2w", "inv")


def read_zkey_header_fflonk(data: bytes) -> dict:
    """src/zkey_utils.js:301-339"""
    secs = read_binfile(data, "zkey", 2)
    if struct.unpack_from("<I", data, secs[1][0])[0] != 10:
        raise SbError("zkey file is not fflonk")                                   # src/fflonk_prove.js:71-73
    p = secs[2][0]
    n8q = struct.unpack_from("<I", data, p)[0]
    q = int.from_bytes(data[p + 4:p + 4 + n8q], "little")
    n8r = struct.unpack_from("<I", data, p + 4 + n8q)[0]
    r = int.from_bytes(data[p + 8 + n8q:p + 8 + n8q + n8r], "little")
    o = p + 8 + n8q + n8r
    n_vars, n_public, domain, n_add, n_cons = struct.unpack_from("<IIIII", data, o)
    return {"protocol": "fflonk", "n8q": n8q, "q": q, "n8r": n8r, "r": r, "nVars": n_vars, "nPublic": n_public,
            "domainSize": domain, "nAdditions": n_add, "nConstraints": n_cons, "power": domain.bit_length() - 1}


def proof_to_object(curve: Curve, raw: bytes) -> dict:
    """Proof.toObjectProof() + stringifyBigInts (src/proof.js:62-82)."""
    n8, q = curve.n8q, curve.q
    pols = {}
    for i, name in enumerate(POINTS):
        p = raw[2 * n8 * i:2 * n8 * (i + 1)]
        pols[name] = ["0", "1", "0"] if p == bytes(2 * n8) else [str(_from_mont(p[:n8], q, n8)), str(_from_mont(p[n8:], q, n8)), "1"]
    base = 2 * n8 * len(POINTS)
    evs = {name: str(_from_mont(raw[base + 32 * i:base + 32 * (i + 1)], curve.r, 32)) for i, name in enumerate(EVALS)}
    return {"polynomials": pols, "evaluations": evs, "protocol": "fflonk", "curve": curve.name}


class ProvingKey:
    """An fflonk zkey resident on one device."""

    def __init__(self, zkey: bytes, curve: Curve | None = None, device: int = 0):
        zkey = bytes(zkey)
        self.header = read_zkey_header_fflonk(zkey)
        self.curve = curve or getCurveFromQ(self.header["q"], device)
        self._own_curve = curve is None
        h = ctypes.c_uint64()
        buf = np.frombuffer(zkey, np.uint8)
        self.curve.check(self.curve.lib.sb_fflonk_load(self.curve.handle, _ptr(buf), buf.size, ctypes.byref(h)))
        self.handle = h.value
        for k in ("nVars", "nPublic", "domainSize", "nAdditions"):
            setattr(self, k, self.header[k])

    @classmethod
    def from_file(cls, path: str, curve: Curve):
        """Loads the key from disk (sb_fflonk_load_file: mapped read-only, streamed to HBM section by section)."""
        self = cls.__new__(cls)
        self.curve, self._own_curve = curve, False
        h = ctypes.c_uint64()
        curve.check(curve.lib.sb_fflonk_load_file(curve.handle, path.encode(), ctypes.byref(h)))   # validates the container
        self.handle = h.value
        nv, npub, ds, na = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        curve.check(curve.lib.sb_fflonk_info(curve.handle, self.handle, ctypes.byref(nv), ctypes.byref(npub), ctypes.byref(ds), ctypes.byref(na)))
        self.nVars, self.nPublic, self.domainSize, self.nAdditions = nv.value, npub.value, ds.value, na.value
        self.header = {"protocol": "fflonk", "r": curve.r, "q": curve.q, "nVars": nv.value, "nPublic": npub.value, "domainSize": ds.value,
                       "nAdditions": na.value, "power": ds.value.bit_length() - 1}
        return self

    def prove_raw(self, witness, blinders: bytes) -> bytes:
        """witness = wtns section 2 ((nVars - nAdditions) x 32 bytes, plain LE), or None to reuse the witness the previous
        proof on this key left in HBM (sb_fflonk_prove_resident); blinders = 9 x 32 Montgomery bytes."""
        if len(blinders) != 9 * 32:
            raise SbError("blinders must be 9 field elements")
        lib, c = self.curve.lib, self.curve
        out = np.empty(lib.sb_fflonk_proof_bytes(c.handle), np.uint8)
        if witness is None:
            c.check(lib.sb_fflonk_prove_resident(c.handle, self.handle, bytes(blinders), _ptr(out)))
        else:
            w = _arr(witness)
            c.check(lib.sb_fflonk_prove(c.handle, self.handle, _ptr(w), w.size // 32, bytes(blinders), _ptr(out)))
        return out.tobytes()

    def release(self):
        if self.handle:
            self.curve.lib.sb_fflonk_release(self.curve.handle, self.handle)
            self.handle = 0
        if self._own_curve:
            self.curve.terminate()


def prove(zkey, wtns: bytes, blinders: bytes | None = None, logger=None, options=None):
    """fflonkProve(zkeyFileName, witnessFileName) -> (proof, publicSignals); zkey may be bytes or a ProvingKey."""
    pk = zkey if isinstance(zkey, ProvingKey) else ProvingKey(zkey)
    try:
        wh, W = read_wtns_header(bytes(wtns))
        if wh["q"] != pk.header["r"]:
            raise SbError("Curve of the witness does not match the curve of the proving key")          # :75-77
        if blinders is None:
            blinders = b"".join(random_fr(pk.curve) for _ in range(9))                                   # :321-324
        raw = pk.prove_raw(np.frombuffer(W, np.uint8), blinders)
        pub = [str(int.from_bytes(W[i * 32:(i + 1) * 32], "little")) for i in range(1, pk.nPublic + 1)]
        return proof_to_object(pk.curve, raw), pub
    finally:
        if not isinstance(zkey, ProvingKey):
            pk.release()
