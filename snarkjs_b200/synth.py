"""Synthetic Groth16 workloads: the chain circuit x_{i+1} = x_i^2 + b with 2^L - 3 constraints
(domain 2^L, nVars = 2^L, nPublic = 1) and a zkey whose base sets are random valid curve points (generated on the
GPU with sb_gen_points, or by any generator with the same definition).  Such a key has no trusted-setup structure, so proofs do not verify, but CPU-vs-GPU parity and
throughput are well defined (every base is a full-size non-trivial point: the worst case for the MSMs)."""
from __future__ import annotations

import struct

import numpy as np

from .curve import Curve, _ptr


def gen_points(curve: Curve, group: int, seed: int, n: int) -> np.ndarray:
    out = np.empty(n * curve.n8q * 2 * group, np.uint8)
    curve.check(curve.lib.sb_gen_points(curve.handle, group, seed, n, _ptr(out)))
    return out


def chain_witness(r: int, L: int, x0: int = 3, b: int = 7) -> np.ndarray:
    """w = [1, out, x0, b, x1 .. x_{N-1}], N = 2^L - 3 constraints, plain little-endian 32-byte values."""
    N = (1 << L) - 3
    xs = [0] * (N + 1)
    xs[0] = x0 % r
    x = xs[0]
    for i in range(N):
        x = (x * x + b) % r
        xs[i + 1] = x
    vals = [1, xs[N], xs[0], b] + xs[1:N]
    buf = bytearray(32 * len(vals))
    for i, v in enumerate(vals):
        buf[32 * i:32 * i + 32] = v.to_bytes(32, "little")
    return np.frombuffer(bytes(buf), np.uint8)


def witness_like(witness: np.ndarray, seed: int = 7) -> np.ndarray:
    """The scalar distribution of real circuits: 50 % zeros, 25 % ones, 25 % untouched; w[0] stays 1."""
    wl = witness.reshape(-1, 32).copy()
    kind = np.random.default_rng(seed).integers(0, 4, wl.shape[0])
    wl[kind < 2] = 0
    wl[kind == 2] = 0
    wl[kind == 2, 0] = 1
    wl[0] = 0
    wl[0, 0] = 1
    return wl.reshape(-1)


def chain_coeffs(r: int, L: int) -> bytes:
    """zkey section 4 for the chain circuit: constraint i is x_i * x_i = x_{i+1} - b  (A and B rows: one entry each),
    plus the nPublic+1 public-input rows appended by setup (src/zkey_new.js:290-300).  Values are 1*R^2 mod r
    (src/zkey_utils.js:174-179)."""
    N = (1 << L) - 3
    nPublic = 1
    R2 = pow(1 << 256, 2, r)
    sig = np.empty(N, np.uint32)            # signal index of x_i: x_0 -> 2, x_i (1 <= i < N) -> 3 + i
    sig[0] = 2
    sig[1:] = 3 + np.arange(1, N, dtype=np.uint32)
    dt = np.dtype([("m", "<u4"), ("c", "<u4"), ("s", "<u4"), ("v", "V32")])
    co = np.zeros(2 * N + nPublic + 1, dt)
    one = np.frombuffer(R2.to_bytes(32, "little"), "V32")[0]
    co["v"] = one
    co["m"][0:2 * N:2] = 0
    co["m"][1:2 * N:2] = 1
    co["c"][0:2 * N:2] = np.arange(N, dtype=np.uint32)
    co["c"][1:2 * N:2] = np.arange(N, dtype=np.uint32)
    co["s"][0:2 * N:2] = sig
    co["s"][1:2 * N:2] = sig
    for s in range(nPublic + 1):
        co[2 * N + s] = (0, N + s, s, one)
    return struct.pack("<I", len(co)) + co.tobytes()


def groth16_zkey_image(q: int, r: int, n8q: int, L: int, gen, seed: int = 1) -> bytes:
    """A Groth16 .zkey image (src/zkey_utils.js:20-45 layout) for the chain circuit with random valid bases.
    gen(group, seed, count) -> affine Montgomery bytes; bench.py passes the GPU generator (sb_gen_points) on the B200 arm
    and the oracle's on the CPU reference arm — both produce the same points, hence byte-identical keys."""
    n8r = 32
    n = 1 << L
    nVars, nPublic = n, 1
    g1 = lambda s, k: bytes(gen(1, seed * 1000003 + s, k))
    g2 = lambda s, k: bytes(gen(2, seed * 1000003 + s, k))
    hdr = struct.pack("<I", n8q) + q.to_bytes(n8q, "little") + struct.pack("<I", n8r) + r.to_bytes(n8r, "little")
    hdr += struct.pack("<III", nVars, nPublic, n)
    hdr += g1(11, 1) + g1(12, 1) + g2(13, 1) + g2(14, 1) + g1(15, 1) + g2(16, 1)     # alpha1 beta1 beta2 gamma2 delta1 delta2
    secs = [
        (1, struct.pack("<I", 1)),
        (2, hdr),
        (3, g1(20, nPublic + 1)),
        (4, chain_coeffs(r, L)),
        (5, g1(1 << 32, nVars)),
        (6, g1(2 << 32, nVars)),
        (7, g2(3 << 32, nVars)),
        (8, g1(4 << 32, nVars - nPublic - 1)),
        (9, g1(5 << 32, n)),
        (10, bytes(64) + struct.pack("<I", 0)),
    ]
    out = [b"zkey" + struct.pack("<II", 1, len(secs))]
    for sid, payload in secs:
        out.append(struct.pack("<IQ", sid, len(payload)))
        out.append(payload)
    return b"".join(out)


def synth_groth16_zkey(curve: Curve, L: int, seed: int = 1) -> bytes:
    """groth16_zkey_image with the bases generated on the GPU."""
    return groth16_zkey_image(curve.q, curve.r, curve.n8q, L, lambda grp, sd, k: gen_points(curve, grp, sd, k).tobytes(), seed)


def wtns_container(r: int, witness: np.ndarray) -> bytes:
    """.wtns image (src/wtns_utils.js:25-37)."""
    nw = witness.size // 32
    s1 = struct.pack("<I", 32) + r.to_bytes(32, "little") + struct.pack("<I", nw)
    return b"wtns" + struct.pack("<II", 2, 2) + struct.pack("<IQ", 1, len(s1)) + s1 + struct.pack("<IQ", 2, witness.size) + witness.tobytes()
