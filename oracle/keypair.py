"""oracle/keypair.py — TEST INFRASTRUCTURE (CPU oracle; never imported by the product).

Restates the hash-to-G2 of the reference's ceremony key pairs, for one purpose: the reference's only hard-coded
known-answer test (test/keypar_test.js:20-119) states pairing equalities  e(g1_sx, g2_sp) == e(g1_s, g2_spx)  in which
g2_sp is *derived* (getG2sp) and everything else is a constant.  Reproducing g2_sp and checking the equalities with the
oracle's own pairing pins, against reference-written numbers: the oracle's BN254 optimal-ate pairing and Fq2 / G2
arithmetic (the code that verifies every Groth16 / PLONK / fflonk proof in tests/), the Montgomery convention of field
elements (fromRng draws the *internal* representation), and the byte conventions of uncompressed points.

  getG2sp      src/keypair.js:38-51    blake2b-512(personalization | challenge | G1.toUncompressed(g1_s) | ...(g1_sx))
  hashToG2     src/keypair.js:24-36    ChaCha seeded with the hash as 8 big-endian words -> G2.fromRng
  ChaCha       ffjavascript, build/snarkjs.js:508-598   (20 rounds, 128-bit block counter, nextU64 = hi word first)
  F1.fromRng   build/snarkjs.js:13019-13031             (n64 little-endian 64-bit draws, masked, rejection; raw = Montgomery)
  G2.fromRng   build/snarkjs.js:13814-13839             (x, `greatest` bit, y = sqrt(x^3 + b) with the sign chosen by
               isNegative, then times the cofactor 16448)
  isNegative   wasmcurves f1m 2950-2963 (fromMontgomery(x) >= (p+1)/2), f2m 4139-4154 (c1 unless c1 == 0, then c0)
"""
from __future__ import annotations

import hashlib

from . import oracle as O

Q = O.P_BN_Q
R_MONT = 1 << 256
B2 = (19485874751759354771024239261021720505790618469301721065564631296452457478373,
      266929791119991161246907387137283842545076965332900288569378510910307636690)     # 3 / (9 + u): build/snarkjs.js pG2b
COFACTOR_G2 = 0x30644e72e131a029b85045b68181585e06ceecda572a2489345f2299c0f9fa8d        # build/snarkjs.js:16448


class ChaCha:
    def __init__(self, seed):
        self.state = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(seed) + [0, 0, 0, 0]
        self.idx, self.buff = 16, [0] * 16

    @staticmethod
    def _qr(st, a, b, c, d):
        rot = lambda v, n: ((v << n) | (v >> (32 - n))) & 0xffffffff
        st[a] = (st[a] + st[b]) & 0xffffffff; st[d] = rot(st[d] ^ st[a], 16)
        st[c] = (st[c] + st[d]) & 0xffffffff; st[b] = rot(st[b] ^ st[c], 12)
        st[a] = (st[a] + st[b]) & 0xffffffff; st[d] = rot(st[d] ^ st[a], 8)
        st[c] = (st[c] + st[d]) & 0xffffffff; st[b] = rot(st[b] ^ st[c], 7)

    def _update(self):
        b = list(self.state)
        for _ in range(10):
            for q in ((0, 4, 8, 12), (1, 5, 9, 13), (2, 6, 10, 14), (3, 7, 11, 15), (0, 5, 10, 15), (1, 6, 11, 12), (2, 7, 8, 13), (3, 4, 9, 14)):
                self._qr(b, *q)
        self.buff = [(x + y) & 0xffffffff for x, y in zip(b, self.state)]
        self.idx = 0
        for i in (12, 13, 14, 15):
            self.state[i] = (self.state[i] + 1) & 0xffffffff
            if self.state[i]:
                break

    def u32(self):
        if self.idx == 16:
            self._update()
        v = self.buff[self.idx]; self.idx += 1
        return v

    def u64(self):
        hi = self.u32()
        return (hi << 32) + self.u32()

    def boolean(self):
        return (self.u32() & 1) == 1


def _f1_from_rng(rng) -> int:
    """-> the element's value (the draw is its Montgomery representation)."""
    mask = (1 << Q.bit_length()) - 1
    while True:
        v = 0
        for i in range(4):
            v += rng.u64() << (64 * i)
        v &= mask
        if v < Q:
            return v * pow(R_MONT, -1, Q) % Q


def _f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
def _f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def _f2_neg(a): return ((-a[0]) % Q, (-a[1]) % Q)


def _f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = _f2_mul(r, a)
        a = _f2_mul(a, a); e >>= 1
    return r


def _f2_sqrt(a):
    """A square root in Fq2 = Fq[u]/(u^2 + 1), q = 3 mod 4, or None (which root does not matter: the caller fixes the sign)."""
    if a == (0, 0):
        return a
    if _f2_pow(a, (Q * Q - 1) // 2) != (1, 0):
        return None
    a1 = _f2_pow(a, (Q - 3) // 4)
    alfa = _f2_mul(_f2_mul(a1, a1), a)
    x0 = _f2_mul(a1, a)
    if alfa == (Q - 1, 0):
        r = _f2_mul((0, 1), x0)
    else:
        r = _f2_mul(_f2_pow(_f2_add((1, 0), alfa), (Q - 1) // 2), x0)
    assert _f2_mul(r, r) == a
    return r


def _f1_is_negative(x): return x >= (Q + 1) // 2
def _f2_is_negative(a): return _f1_is_negative(a[0]) if a[1] == 0 else _f1_is_negative(a[1])


def g2_from_rng(rng):
    while True:
        x = (_f1_from_rng(rng), _f1_from_rng(rng))
        greatest = rng.boolean()
        x3b = _f2_add(_f2_mul(_f2_mul(x, x), x), B2)
        y = _f2_sqrt(x3b)
        if y is not None:
            break
    if greatest ^ _f2_is_negative(y):
        y = _f2_neg(y)
    ci = O.CURVES[O.BN254]
    jac = O.g_from_affine(O.BN254, 2, ci.g2_affine_bytes((x, y)))
    out = O.g_to_affine(O.BN254, 2, O.g_times(O.BN254, 2, jac, COFACTOR_G2.to_bytes(32, "little")))
    return ci.g2_from_affine_bytes(out)


def g1_uncompressed(pt) -> bytes:
    """G1.toUncompressed: x | y, 32 big-endian bytes each, out of Montgomery form (wasmcurves _LEMtoU 7123-7148)."""
    return pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


def get_g2sp(personalization: int, challenge: bytes, g1_s, g1_sx):
    h = hashlib.blake2b(digest_size=64)
    h.update(bytes([personalization])); h.update(challenge); h.update(g1_uncompressed(g1_s)); h.update(g1_uncompressed(g1_sx))
    d = h.digest()
    return g2_from_rng(ChaCha([int.from_bytes(d[4 * i:4 * i + 4], "big") for i in range(8)]))
