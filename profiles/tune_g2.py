"""Times the G2 bucket-accumulation kernel variants (sb_set_tuning(0, v)) on a registered 2^20 G2 MSM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snarkjs_b200
from snarkjs_b200 import synth
c = snarkjs_b200.getCurveFromName("bn128")
lib, h = c.lib, c.handle
n = 1 << 20
rng = np.random.default_rng(1)
sc = rng.integers(0, 256, size=n * 32, dtype=np.uint8); sc.reshape(n, 32)[:, 31] &= 0x1f
for grp in (2, 1):
    bases = synth.gen_points(c, grp, 7, n)
    G = c.G1 if grp == 1 else c.G2
    hb = G.registerBases(bases)
    ref = None
    for v in ((3, 2, 4, 3) if grp == 2 else (4,)):
        lib.sb_set_tuning(0, v)
        for _ in range(3):
            out = G.multiExpRegistered(hb, sc)
        if ref is None: ref = out.tobytes()
        assert out.tobytes() == ref
        print(f"G{grp} variant {v}: acc kernel {lib.sb_last_stat(h, grp - 1):.3f} ms ({lib.sb_last_stat(h, 3 + grp):.0f} entries), msm total {c.last_ms(0):.3f} ms (h2d {c.last_ms(1):.3f})")
