#!/usr/bin/env python
"""Writes tests/golden/bench_proof_hashes.json: sha256 of the CPU oracle's Groth16 proof object on bench.py's synthetic
chain-circuit key (snarkjs_b200/synth.py, bases from the oracle's generator, r = 5, s = 7) for the domain sizes the bench
and the tests use.  bench.py asserts its GPU proof against these on every run, so the timed workload is parity-checked
at the benchmark size (2^20) and at config #4's size (2^22), not only in the 2^16 pytest case.

    python tests/golden/make_bench_hashes.py 12 16 18 20 22      (minutes on 8 cores; needs no GPU)
    python tests/golden/make_bench_hashes.py plonk:bls12381:16 plonk:bls12381:20 fflonk:bn128:20
        the PLONK / fflonk bench workloads (bench_plonk.py): proof of the CPU port (tests/host/ flow with OpenMP, oracle
        NTT / MSM — itself checked proof for proof against oracle/plonk.py and oracle/fflonk.py in tests/test_host_*.py)
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "bench_proof_hashes.json")


def proof_hash(proof_obj) -> str:
    return hashlib.sha256(json.dumps(proof_obj, sort_keys=True, separators=(",", ":")).encode()).hexdigest()


def oracle_bench_proof(L: int, witness_like: bool = False):
    from oracle import oracle as O
    from snarkjs_b200 import synth
    ci = O.CURVES[O.BN254]
    zkey = synth.groth16_zkey_image(ci.q, ci.r, 32, L, lambda g, s, k: O.gen_points(O.BN254, g, s, k).tobytes())
    wit = synth.chain_witness(ci.r, L)
    if witness_like:
        wit = synth.witness_like(wit)
    wt = synth.wtns_container(ci.r, wit)
    proof, _ = O.groth16_prove(zkey, wt, ci.fr_to_mont(5), ci.fr_to_mont(7), concurrency=max(1, O.lib().or_num_threads()))
    return proof


if __name__ == "__main__":
    table = json.load(open(OUT)) if os.path.exists(OUT) else {}
    tab = table.setdefault("groth16_bn128_chain_r5_s7", {})
    for a in sys.argv[1:]:
        if ":" in a:
            from types import SimpleNamespace
            import bench_plonk as B
            proto, cname, L = a.split(":")
            t = time.time()
            zkey, wit, ci = B.oracle_key(proto, cname, int(L))
            _, raw = B.cpu_prove(proto, zkey, wit, ci.r, ci.n8q, os.cpu_count())
            ns = SimpleNamespace(name=cname, n8q=ci.n8q, q=ci.q, r=ci.r)
            h = proof_hash(B._proof_object(proto, ns, raw))
            table.setdefault(f"{proto}_{cname}_chain_b7", {})[L] = h
            print(a, h, f"{time.time() - t:.1f}s", flush=True)
            json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
            continue
        L = int(a)
        t = time.time()
        tab[str(L)] = proof_hash(oracle_bench_proof(L))
        print(L, tab[str(L)], f"{time.time() - t:.1f}s", flush=True)
        json.dump(table, open(OUT, "w"), indent=1, sort_keys=True)
