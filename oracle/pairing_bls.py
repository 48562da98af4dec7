"""oracle/pairing_bls.py — BLS12-381 ate pairing over plain Python ints.  TEST INFRASTRUCTURE ONLY (verifiers, small inputs).

Same construction as the BN254 pairing in oracle.py (restating what ffjavascript's `pairingEq` decides,
build/snarkjs.js:10689-10780 for BN254, 12184-12260 for BLS12-381): Fq12 = Fq[w]/(w^12 - 2 w^6 + 2) with
Fq2 = Fq[u]/(u^2 + 1) embedded by u = w^6 - 1, the M-type twist (x, y) -> (x / w^2, y / w^3), a Miller loop over
|x| = 0xd201000000010000 without Frobenius corrections, and the plain final exponentiation f^((q^12 - 1) / r).
Only products of pairings compared with 1 are exposed, so the sign convention of x does not matter.
Checked in tests/test_oracle_plonk.py: bilinearity on the generators, and PLONK proofs on BLS12-381 keys verify."""
from __future__ import annotations

Q = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
ATE = 0xd201000000010000
LOG_ATE = 63


def _mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for i in range(22, 11, -1):           # w^12 = 2 w^6 - 2
        v = t[i]
        if v:
            t[i - 6] += 2 * v
            t[i - 12] -= 2 * v
    return [x % Q for x in t[:12]]


ONE = [1] + [0] * 11


def _add(a, b): return [(x + y) % Q for x, y in zip(a, b)]
def _sub(a, b): return [(x - y) % Q for x, y in zip(a, b)]
def _scale(a, k): return [x * k % Q for x in a]


def _pow(a, e):
    r, b = ONE, a
    while e:
        if e & 1:
            r = _mul(r, b)
        b = _mul(b, b)
        e >>= 1
    return r


def _inv(a):
    """Inverse in Fq12 by the extended Euclidean algorithm over Fq[x]."""
    mod = [2, 0, 0, 0, 0, 0, Q - 2, 0, 0, 0, 0, 0, 1]
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], mod

    def deg(p):
        d = len(p) - 1
        while d and p[d] == 0:
            d -= 1
        return d

    def pdiv(x, y):
        dx, dy = deg(x), deg(y)
        t, o = list(x), [0] * len(x)
        iy = pow(y[dy], -1, Q)
        for i in range(dx - dy, -1, -1):
            o[i] = (o[i] + t[dy + i] * iy) % Q
            for c in range(dy + 1):
                t[c + i] = (t[c + i] - o[i] * y[c]) % Q
        return o[:deg(o) + 1]

    while deg(low):
        r = pdiv(high, low)
        r += [0] * (13 - len(r))
        nm, new = list(hm), list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] -= lm[i] * r[j]
                new[i + j] -= low[i] * r[j]
        nm = [x % Q for x in nm]
        new = [x % Q for x in new]
        lm, low, hm, high = nm, new, lm, low
    il = pow(low[0], -1, Q)
    return [x * il % Q for x in lm[:12]]


_W = [0, 1] + [0] * 10
_W2I = _inv(_mul(_W, _W))
_W3I = _inv(_mul(_mul(_W, _W), _W))


def _twist(pt):
    (x0, x1), (y0, y1) = pt
    nx = [(x0 - x1) % Q] + [0] * 5 + [x1] + [0] * 5        # u = w^6 - 1
    ny = [(y0 - y1) % Q] + [0] * 5 + [y1] + [0] * 5
    return (_mul(nx, _W2I), _mul(ny, _W3I))


def _double(p):
    x, y = p
    m = _mul(_scale(_mul(x, x), 3), _inv(_scale(y, 2)))
    nx = _sub(_mul(m, m), _scale(x, 2))
    return (nx, _sub(_mul(m, _sub(x, nx)), y))


def _addp(p, q):
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        return _double(p) if y1 == y2 else None
    m = _mul(_sub(y2, y1), _inv(_sub(x2, x1)))
    nx = _sub(_sub(_mul(m, m), x1), x2)
    return (nx, _sub(_mul(m, _sub(x1, nx)), y1))


def _line(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if x1 != x2:
        m = _mul(_sub(y2, y1), _inv(_sub(x2, x1)))
        return _sub(_mul(m, _sub(xt, x1)), _sub(yt, y1))
    if y1 == y2:
        m = _mul(_scale(_mul(x1, x1), 3), _inv(_scale(y1, 2)))
        return _sub(_mul(m, _sub(xt, x1)), _sub(yt, y1))
    return _sub(xt, x1)


def _miller(q2, p1):
    if q2 is None or p1 is None:
        return ONE
    Qt = _twist(q2)
    P = ([p1[0] % Q] + [0] * 11, [p1[1] % Q] + [0] * 11)
    Rr, f = Qt, ONE
    for i in range(LOG_ATE - 1, -1, -1):
        f = _mul(_mul(f, f), _line(Rr, Rr, P))
        Rr = _double(Rr)
        if ATE & (1 << i):
            f = _mul(f, _line(Rr, Qt, P))
            Rr = _addp(Rr, Qt)
    return f


def pairing_product_is_one(pairs) -> bool:
    """prod e(P_i, Q_i) == 1 for (G1 affine ints, G2 affine ((x0, x1), (y0, y1)) ints) pairs on BLS12-381."""
    f = ONE
    for p1, q2 in pairs:
        f = _mul(f, _miller(q2, p1))
    return _pow(f, (Q ** 12 - 1) // R) == ONE


# G1 arithmetic on affine int pairs (None = infinity); y^2 = x^3 + 4
def g1_valid(pt) -> bool:
    return pt is None or (pt[1] * pt[1] - pt[0] ** 3 - 4) % Q == 0


def g1_neg(pt):
    return None if pt is None else (pt[0], (-pt[1]) % Q)


def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % Q == 0:
            return None
        m = 3 * p[0] * p[0] * pow(2 * p[1], -1, Q) % Q
    else:
        m = (q[1] - p[1]) * pow(q[0] - p[0], -1, Q) % Q
    x = (m * m - p[0] - q[0]) % Q
    return (x, (m * (p[0] - x) - p[1]) % Q)


def g1_mul(p, k):
    r = None
    while k:
        if k & 1:
            r = g1_add(r, p)
        p = g1_add(p, p)
        k >>= 1
    return r
