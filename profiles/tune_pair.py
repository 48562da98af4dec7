"""MSM time on registered 2^20 G1/G2 bases for the pairing-round settings: sb_set_tuning(4, off) / sb_set_tuning(5, max rounds)."""
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
for grp in (1, 2):
    bases = synth.gen_points(c, grp, 7, n)
    G = c.G1 if grp == 1 else c.G2
    hb = G.registerBases(bases)
    ref = None
    for off, rounds in ((1, 0), (0, 1), (0, 2), (0, 3), (0, 4), (0, 6)):
        lib.sb_set_tuning(4, off); lib.sb_set_tuning(5, rounds)
        for _ in range(3):
            out = G.multiExpRegistered(hb, sc)
        if ref is None: ref = out.tobytes()
        assert out.tobytes() == ref
        print(f"G{grp} pairing {'off' if off else 'R<=%d' % rounds}: pre-stage {lib.sb_last_stat(h, grp - 1):.3f} ms, msm device total {c.last_ms(2):.3f} ms")
