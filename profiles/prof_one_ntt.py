"""BASELINE config #3: one forward + one inverse Fr NTT of 2^NTT_LOGN elements (BN254) on device-resident data inside a
cudaProfilerStart/Stop window (ncu --profile-from-start off), plus CUDA-event timings without the profiler's help."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import snarkjs_b200
from snarkjs_b200.curve import _ptr
L = int(os.environ.get("NTT_LOGN", "24"))
c = snarkjs_b200.getCurveFromName("bn128")
lib, h = c.lib, c.handle
m = 1 << L
rng = np.random.default_rng(4)
x = rng.integers(0, 256, size=m * 32, dtype=np.uint8); x.reshape(m, 32)[:, 31] &= 0x1f
lib.sb_dev_alloc.restype = ctypes.c_void_p
a = ctypes.c_void_p(lib.sb_dev_alloc(h, m * 32)); b = ctypes.c_void_p(lib.sb_dev_alloc(h, m * 32))
lib.sb_dev_upload(h, a, _ptr(x), m * 32)
res = ctypes.c_void_p()
def rt():
    c.check(lib.sb_ntt_fr_dev(h, a, b, m, 0, ctypes.byref(res)))
    src = res.value; other = b.value if src == a.value else a.value
    c.check(lib.sb_ntt_fr_dev(h, ctypes.c_void_p(src), ctypes.c_void_p(other), m, 1, ctypes.byref(res)))
for _ in range(2):
    rt()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
t = time.perf_counter(); rt(); lib.sb_sync(h); dt = time.perf_counter() - t
torch.cuda.profiler.stop()
print("ntt+intt 2^%d: %.3f ms wall (includes two launches' host overhead)" % (L, dt * 1e3))
