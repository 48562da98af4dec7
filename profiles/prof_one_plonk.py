"""One PLONK (default, BLS12-381) or fflonk (PROTO=fflonk, BN254) proof at domain 2^LOGN inside a cudaProfilerStart/Stop
window, for `ncu --profile-from-start off`.  The key is bench.py's synthetic chain circuit (snarkjs_b200/synth.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import snarkjs_b200
from snarkjs_b200 import synth
import bench_plonk as B
L = int(os.environ.get("LOGN", "16"))
proto = os.environ.get("PROTO", "plonk")
cname = os.environ.get("CURVE", "bls12381" if proto == "plonk" else "bn128")
c = snarkjs_b200.getCurveFromName(cname)
zkey, W = (synth.synth_plonk_zkey if proto == "plonk" else synth.synth_fflonk_zkey)(c, L)
mod = snarkjs_b200.plonk if proto == "plonk" else snarkjs_b200.fflonk
pk = mod.ProvingKey(zkey, c)
del zkey
bl = B._blinders(c.r, proto)
for _ in range(2):
    pk.prove_raw(W, bl)
torch.cuda.synchronize()
torch.cuda.profiler.start()
pk.prove_raw(W, bl)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", proto, cname, L, c.last_ms(0), [round(c.last_ms(i), 3) for i in range(1, 6)])
