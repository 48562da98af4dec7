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


# ------------------------------------------------------------------------------------------------ PLONK / fflonk keys
# Synthetic PLONK-shaped keys for BASELINE.json config #5 (plonk prove + fflonk prove at domain 2^20).  The circuit is
# given directly as gates (the reference derives them from an r1cs, src/plonk_setup.js:142-299): the chain
# x_{i+1} = x_i^2 + c with public output, plus two levels of linear "addition" wires (the shape reduceCoefs emits,
# plonk_setup.js:176-215).  Everything per-element is numpy; the field work (the [coef n | evals 4n] blocks of
# plonk_setup.js:331-338 and the powers of omega) goes through two callables, so that bench.py builds the key with the
# library's own NTT on the B200 arm and with the CPU oracle's on the reference arm: both give the same bytes.
#   fft(buf uint8[n*32], inverse) -> uint8[n*32]          (Montgomery in/out, natural order)
#   apply_key(buf, first32, inc32) -> buf[i]*first*inc^i  (batchApplyKey)
#   gen(group, seed, count) -> affine Montgomery point bytes
# The PTau section holds pseudo-random valid points (no structure: such proofs do not verify, parity and throughput are
# well defined); tests/test_py_mirror.py checks these builders byte for byte against the oracle's restated setup.

def _mont(x: int, r: int) -> bytes:
    return ((x << 256) % r).to_bytes(32, "little")


def _rows(ints, r: int) -> np.ndarray:
    return np.frombuffer(b"".join(_mont(int(v) % r, r) for v in ints), np.uint8).reshape(-1, 32)


def plonk_chain_circuit(n_gates: int, r: int, seed: int = 7) -> dict:
    """Gate arrays of the chain circuit (same definition as the oracle's chain_gates with one public signal)."""
    n_pub = 1
    n_y = max(2, n_gates // 8)
    n_z = n_y - 1
    m = n_gates - n_pub - n_y - n_z
    assert m >= n_y + 1
    cst = (seed * 0x9E3779B97F4A7C15 + 12345) % r
    x = [0] * (m + 1)
    x[0] = (seed * 1000003 + 17) % r
    v = x[0]
    for i in range(m):
        v = (v * v + cst) % r
        x[i + 1] = v
    wire_x = np.empty(m + 1, np.uint32)
    wire_x[:m] = 2 + np.arange(m, dtype=np.uint32)
    wire_x[m] = 1
    n_wit = m + 2
    wire_y = n_wit + np.arange(n_y, dtype=np.uint32)
    wire_z = n_wit + n_y + np.arange(n_z, dtype=np.uint32)
    sl = np.concatenate([np.array([1], np.uint32), wire_x[:m], wire_x[:n_y], wire_y[:n_z]]).astype(np.uint32)
    sr = np.concatenate([np.array([0], np.uint32), wire_x[:m], wire_x[1:n_y + 1], wire_y[1:n_z + 1]]).astype(np.uint32)
    so = np.concatenate([np.array([0], np.uint32), wire_x[1:m + 1], wire_y, wire_z]).astype(np.uint32)
    # selector values as codes into a small constant table
    table = [0, 1, r - 1, r - 3, r - 7, r - 2, cst]
    seg = [n_pub, m, n_y, n_z]

    def codes(a, b, c, d):
        return np.repeat(np.array([a, b, c, d], np.uint8), seg)
    sel = {"qm": codes(0, 1, 0, 0), "ql": codes(1, 0, 3, 2), "qr": codes(0, 0, 4, 5), "qo": codes(0, 2, 1, 1), "qc": codes(0, 6, 0, 0)}
    add = np.zeros(n_y + n_z, np.dtype([("a", "<u4"), ("b", "<u4"), ("f1", "V32"), ("f2", "V32")]))
    add["a"][:n_y], add["b"][:n_y] = wire_x[:n_y], wire_x[1:n_y + 1]
    add["a"][n_y:], add["b"][n_y:] = wire_y[:n_z], wire_y[1:n_z + 1]
    f = {k: np.frombuffer(_mont(k, r), "V32")[0] for k in (1, 2, 3, 7)}
    add["f1"][:n_y], add["f2"][:n_y], add["f1"][n_y:], add["f2"][n_y:] = f[3], f[7], f[1], f[2]
    wit = [1, x[m]] + x[:m]
    witness = np.frombuffer(b"".join(int(w).to_bytes(32, "little") for w in wit), np.uint8)
    return {"sl": sl, "sr": sr, "so": so, "sel": sel, "table": table, "additions": add, "n_vars": n_wit + n_y + n_z, "n_public": n_pub,
            "n_gates": n_gates, "witness": witness}


def _k1k2(n: int, r: int):
    k1 = 2                                                   # getK1K2, src/plonk_setup.js:482-510
    while pow(k1, n, r) == 1:
        k1 += 1
    k2 = k1 + 1
    while pow(k2, n, r) == 1 or pow(k2 * pow(k1, -1, r) % r, n, r) == 1:
        k2 += 1
    return k1, k2


def _sigma_rows(circ: dict, n: int, r: int, wn: int, k1: int, k2: int, apply_key, free_rows: int) -> np.ndarray:
    """sigma evaluations (3n x 32 Montgomery bytes, position p = col*n + i), src/plonk_setup.js:362-438: every wire's
    occurrences, in visiting order (row-major), are mapped cyclically onto each other.  free_rows > 0: the last rows map
    to themselves (fflonk's two blinding rows, src/fflonk_setup.js:356-360)."""
    ng = circ["n_gates"]
    one = np.tile(np.frombuffer(_mont(1, r), np.uint8), n)
    V = np.concatenate([np.asarray(apply_key(one, _mont(k, r), _mont(wn, r))).reshape(n, 32) for k in (1, k1, k2)])
    wires = np.zeros((n, 3), np.uint32)
    wires[:ng, 0], wires[:ng, 1], wires[:ng, 2] = circ["sl"], circ["sr"], circ["so"]
    live = n - free_rows
    flat = wires[:live].reshape(-1)
    t = np.arange(flat.size, dtype=np.int64)
    pos = (t % 3) * n + t // 3
    order = np.argsort(flat, kind="stable")
    sw = flat[order]
    start = np.ones(sw.size, bool)
    start[1:] = sw[1:] != sw[:-1]
    end = np.ones(sw.size, bool)
    end[:-1] = start[1:]
    gid = np.cumsum(start) - 1
    last_of_group = np.nonzero(end)[0][gid]                   # sorted index of the group's last element
    prev_sorted = np.where(start, last_of_group, np.arange(sw.size) - 1)
    src = np.empty(flat.size, np.int64)
    src[order] = order[prev_sorted]
    sigma = V.copy()                                          # free rows (and nothing else) keep their own value
    sigma[pos] = V[pos[src]]
    return sigma


def _p4(evals: np.ndarray, n: int, fft):
    coef = np.asarray(fft(np.ascontiguousarray(evals).reshape(-1), True)).reshape(-1)
    ev4 = np.asarray(fft(np.concatenate([coef, np.zeros(3 * n * 32, np.uint8)]), False)).reshape(-1)
    return coef, ev4


def _binfile(kind: bytes, version: int, secs) -> bytes:
    out = [kind + struct.pack("<II", version, len(secs))]
    for sid, payload in secs:
        parts = payload if isinstance(payload, (list, tuple)) else [payload]
        parts = [p.tobytes() if isinstance(p, np.ndarray) else bytes(p) for p in parts]
        out.append(struct.pack("<IQ", sid, sum(len(p) for p in parts)))
        out.extend(parts)
    return b"".join(out)


def _selector_evals(circ: dict, name: str, n: int, r: int) -> np.ndarray:
    rows = _rows(circ["table"], r)
    ev = np.zeros((n, 32), np.uint8)
    ev[:circ["n_gates"]] = rows[circ["sel"][name]]
    return ev


def plonk_zkey_image(q: int, r: int, n8q: int, circ: dict, wn_of, fft, apply_key, gen, g2_generator: bytes, seed: int = 4242) -> bytes:
    """A PLONK .zkey image (sections of src/plonk_setup.js:99-480, header src/zkey_utils.js:268-294) for `circ`.
    wn_of(power) -> the reference's 2^power-th root of unity as an int."""
    ng = circ["n_gates"]
    power = max(3, (ng - 1).bit_length())                     # plonk_setup.js:74-76
    n = 1 << power
    wn = wn_of(power)
    k1, k2 = _k1k2(n, r)
    secs = [(3, circ["additions"].tobytes()), (4, circ["sl"].astype("<u4")), (5, circ["sr"].astype("<u4")), (6, circ["so"].astype("<u4"))]
    for sid, name in ((7, "qm"), (8, "ql"), (9, "qr"), (10, "qo"), (11, "qc")):
        secs.append((sid, list(_p4(_selector_evals(circ, name, n, r), n, fft))))
    sigma = _sigma_rows(circ, n, r, wn, k1, k2, apply_key, 0)
    s12 = []
    for col in range(3):
        s12.extend(_p4(sigma[col * n:(col + 1) * n], n, fft))
    secs.append((12, s12))
    lag = []
    for i in range(max(circ["n_public"], 1)):                 # writeLs, plonk_setup.js:440-450
        ev = np.zeros((n, 32), np.uint8)
        ev[i] = np.frombuffer(_mont(1, r), np.uint8)
        lag.extend(_p4(ev, n, fft))
    secs.append((13, lag))
    pts = bytes(gen(1, seed & 0xFFFFFFFF, n + 6))
    secs.append((14, pts))
    sg = 2 * n8q
    hdr = struct.pack("<I", n8q) + q.to_bytes(n8q, "little") + struct.pack("<I", 32) + r.to_bytes(32, "little")
    hdr += struct.pack("<IIIII", circ["n_vars"], circ["n_public"], n, len(circ["additions"]), ng)
    hdr += _mont(k1, r) + _mont(k2, r)
    for i in range(7, -1, -1):                                # Qm Ql Qr Qo Qc S1 S2 S3: any valid points (only hashed)
        hdr += pts[i * sg:(i + 1) * sg]
    hdr += bytes(g2_generator)
    return _binfile(b"zkey", 1, [(1, struct.pack("<I", 2)), (2, hdr)] + secs)


def fflonk_zkey_image(q: int, r: int, n8q: int, circ: dict, wn_of, fft, apply_key, gen, g2_generator: bytes, seed: int = 4242) -> bytes:
    """A fflonk .zkey image (src/fflonk_setup.js:211-503; BN254 only like the reference's w3 / wr constants :534-557)."""
    assert n8q == 32, "fflonk: BN254 only"
    ng = circ["n_gates"]
    power = max(3, (ng + 2 - 1).bit_length())                 # fflonk_setup.js:112 (two rows stay free for blinding)
    n = 1 << power
    wn = wn_of(power)
    k1, k2 = _k1k2(n, r)
    w3 = pow(31624, 3648040478639879203707734290876212514758060733402672390616367364429301415936 // 3, r)   # computeW3 :534-542
    w4, w8 = wn_of(2), wn_of(3)
    wr = pow(467799165886069610036046866799264026481344299079011762026774533774345988080, 1 << (28 - power), r)
    secs = [(3, circ["additions"].tobytes()), (4, circ["sl"].astype("<u4")), (5, circ["sr"].astype("<u4")), (6, circ["so"].astype("<u4"))]
    coefs = {}
    for sid, name, key in ((7, "QL", "ql"), (8, "QR", "qr"), (9, "QM", "qm"), (10, "QO", "qo"), (11, "QC", "qc")):
        c, e = _p4(_selector_evals(circ, key, n, r), n, fft)
        coefs[name] = c
        secs.append((sid, [c, e]))
    sigma = _sigma_rows(circ, n, r, wn, k1, k2, apply_key, 2)
    for col, name in enumerate(("S1", "S2", "S3")):
        c, e = _p4(sigma[col * n:(col + 1) * n], n, fft)
        coefs[name] = c
        secs.append((12 + col, [c, e]))
    lag = []
    for i in range(max(circ["n_public"], 1)):
        ev = np.zeros((n, 32), np.uint8)
        ev[i] = np.frombuffer(_mont(1, r), np.uint8)
        lag.extend(_p4(ev, n, fft))
    secs.append((15, lag))
    pts = bytes(gen(1, seed & 0xFFFFFFFF, 9 * n + 18))
    secs.append((16, pts))
    c0 = np.stack([coefs[k].reshape(n, 32) for k in ("QL", "QR", "QO", "QM", "QC", "S1", "S2", "S3")], axis=1)   # writeC0 :441-464
    secs.append((17, c0.reshape(-1)))
    hdr = struct.pack("<I", n8q) + q.to_bytes(n8q, "little") + struct.pack("<I", 32) + r.to_bytes(32, "little")
    hdr += struct.pack("<IIIII", circ["n_vars"], circ["n_public"], n, len(circ["additions"]), ng)
    for v in (k1, k2, w3, w4, w8, wr):
        hdr += _mont(v, r)
    hdr += bytes(g2_generator) + pts[:2 * n8q]
    return _binfile(b"zkey", 1, [(1, struct.pack("<I", 10)), (2, hdr)] + secs)


def curve_callbacks(curve: Curve):
    """(wn_of, fft, apply_key, gen, g2_generator) backed by the library (GPU)."""
    g2 = np.empty(4 * curve.n8q, np.uint8)
    curve.check(curve.lib.sb_generator(curve.handle, 2, _ptr(g2)))
    R_inv = pow(1 << 256, -1, curve.r)
    wn_of = lambda p: int.from_bytes(curve.Fr.w[p], "little") * R_inv % curve.r
    return (wn_of, lambda b, inv: curve.Fr.ifft(b) if inv else curve.Fr.fft(b), curve.Fr.batchApplyKey,
            lambda grp, sd, k: gen_points(curve, grp, sd, k), g2.tobytes())


def synth_plonk_zkey(curve: Curve, log_n: int, seed: int = 4242):
    """(zkey image, witness section bytes) of the chain circuit with 2^log_n - 6 gates, built with the library's NTT."""
    circ = plonk_chain_circuit((1 << log_n) - 6, curve.r)
    return plonk_zkey_image(curve.q, curve.r, curve.n8q, circ, *curve_callbacks(curve), seed=seed), circ["witness"]


def synth_fflonk_zkey(curve: Curve, log_n: int, seed: int = 4242):
    circ = plonk_chain_circuit((1 << log_n) - 6, curve.r)
    return fflonk_zkey_image(curve.q, curve.r, curve.n8q, circ, *curve_callbacks(curve), seed=seed), circ["witness"]
