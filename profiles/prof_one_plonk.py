"""One PLONK (default) or fflonk (PROTO=fflonk) proof at domain 2^LOGN inside a cudaProfilerStart/Stop window, for
`ncu --profile-from-start off`.  The key is a synthetic satisfiable circuit on unstructured points (oracle setup, CPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("OMP_NUM_THREADS", "8")
import numpy as np, torch
import snarkjs_b200
from oracle import oracle as orc, plonk as op, fflonk as off
L = int(os.environ.get("LOGN", "16"))
proto = os.environ.get("PROTO", "plonk")
c = snarkjs_b200.getCurveFromName("bn128")
ci = orc.CURVES[orc.BN254]
gates, adds, n_vars, n_pub, wit = op.chain_gates((1 << L) - 6)
setup = op.plonk_setup_synth if proto == "plonk" else off.fflonk_setup_synth
mod = snarkjs_b200.plonk if proto == "plonk" else snarkjs_b200.fflonk
pk = mod.ProvingKey(setup(gates, adds, n_vars, n_pub, tau=4242, structured=False), c)
W = np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in wit), np.uint8)
bl = b"".join(ci.fr_to_mont(7 + i) for i in range(11 if proto == "plonk" else 9))
for _ in range(2):
    pk.prove_raw(W, bl)
torch.cuda.synchronize()
torch.cuda.profiler.start()
pk.prove_raw(W, bl)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", proto, L, c.last_ms(0), [round(c.last_ms(i), 3) for i in range(1, 6)])
