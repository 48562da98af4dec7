"""GPU: the bench workloads of BASELINE.json config #5 at a size the CPU port finishes in seconds — PLONK on BLS12-381 and
fflonk on BN254 at domain 2^16 (the verdict's ">= 2^18" cases run in bench.py --workload plonk|fflonk at 2^20 with the same
live comparison).  The key is built by snarkjs_b200/synth.py with the library's NTT; the CPU port (tests/host/ flow compiled
with OpenMP, oracle NTT / MSM) proves the key built with the oracle's NTT.  Both keys must be the same bytes and both proofs
the same bytes; the resident-witness entry must reproduce the proof."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("proto,cname,L", [("plonk", "bls12381", 16), ("fflonk", "bn128", 16), ("plonk", "bn128", 13)])
def test_bench_workload_matches_cpu_port(proto, cname, L):
    import snarkjs_b200
    from snarkjs_b200 import fflonk, plonk, synth
    import bench_plonk as B
    curve = snarkjs_b200.getCurveFromName(cname)
    try:
        zkey, wit = (synth.synth_plonk_zkey if proto == "plonk" else synth.synth_fflonk_zkey)(curve, L)
        ozkey, owit, ci = B.oracle_key(proto, cname, L)
        assert hashlib.sha256(zkey).digest() == hashlib.sha256(ozkey).digest(), "GPU-built and oracle-built keys differ"
        assert wit.tobytes() == owit.tobytes()
        mod = plonk if proto == "plonk" else fflonk
        pk = mod.ProvingKey(zkey, curve)
        bl = B._blinders(curve.r, proto)
        raw = pk.prove_raw(wit, bl)
        _, want = B.cpu_prove(proto, ozkey, owit, ci.r, ci.n8q, 8)
        assert raw == want
        assert pk.prove_raw(None, bl) == raw                       # witness resident in HBM
        bl2 = bl[32:] + bl[:32]
        assert pk.prove_raw(None, bl2) == pk.prove_raw(wit, bl2) != raw
        pk.release()
    finally:
        curve.terminate()


def test_resident_needs_a_witness():
    import snarkjs_b200
    from snarkjs_b200 import plonk, synth
    from snarkjs_b200.curve import SbError
    import bench_plonk as B
    curve = snarkjs_b200.getCurveFromName("bn128")
    try:
        zkey, wit = synth.synth_plonk_zkey(curve, 8)
        pk = plonk.ProvingKey(zkey, curve)
        with pytest.raises(SbError, match="no witness resident"):
            pk.prove_raw(None, B._blinders(curve.r, "plonk"))
        pk.release()
    finally:
        curve.terminate()
