"""One Groth16 proof (domain 2^LOGN) inside a cudaProfilerStart/Stop window, for ncu --profile-from-start off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import snarkjs_b200
from snarkjs_b200 import groth16, synth
from snarkjs_b200.curve import _ptr
L = int(os.environ.get("LOGN", "20"))
c = snarkjs_b200.getCurveFromName("bn128")
if os.environ.get("SB_SERIAL"): c.lib.sb_set_tuning(2, 1)
if os.environ.get("SB_CALIB"): print("calib", c.lib.sb_calibrate(c.handle, 1)); sys.exit(0)
pk = groth16.ProvingKey(synth.synth_groth16_zkey(c, L, seed=1), curve=c)
w = synth.chain_witness(c.r, L)
r = s = (5 * (1 << 256) % c.r).to_bytes(32, "little")
proof = np.empty(256, np.uint8)
for _ in range(2):
    c.check(c.lib.sb_groth16_prove(c.handle, pk.handle, _ptr(w), w.size // 32, r, s, _ptr(proof)))
torch.cuda.synchronize()
torch.cuda.profiler.start()
c.check(c.lib.sb_groth16_prove(c.handle, pk.handle, _ptr(w), w.size // 32, r, s, _ptr(proof)))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", c.last_ms(0))
