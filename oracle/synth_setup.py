"""oracle/synth_setup.py — synthetic inputs for `zkey new` with KNOWN toxic waste.  TEST INFRASTRUCTURE ONLY.

The reference ships one prepared powers-of-tau file, on BN254.  To exercise the Groth16 path on BLS12-381 with keys whose
proofs *verify* (not only compare), this module writes, for any curve:
  * chain_r1cs(...)        an .r1cs container (r1csfile layout) for the chain x_{i+1} = x_i^2 + b with one
                           public output, plus its witness — the shape of test/groth16/circuit.circom;
  * prepared_ptau(...)     the sections of a prepared .ptau that src/zkey_new.js:101-151,182-200 reads (4 alphaTauG1,
                           5 betaTauG1, 6 betaG2, and the Lagrange-basis sections 12-15 for one domain size and its double),
                           computed from chosen tau, alpha, beta instead of a ceremony.
oracle.zkey_new(r1cs_bytes, ptau_bytes) then builds the zkey exactly as for the reference's files."""
from __future__ import annotations

import struct
from typing import List, Tuple

from . import oracle as orc


def chain_r1cs(curve: int, n_constraints: int, seed: int = 5) -> Tuple[bytes, List[int]]:
    """Wires: 0 = one, 1 = x_m (public output), 2 = x_0, 3.. = x_1..x_{m-1}; constraint i: x_i * x_i = x_{i+1} - b."""
    ci = orc.CURVES[curve]
    r = ci.r
    m = n_constraints
    b = (seed * 7919 + 3) % r
    x = [(seed * 104729 + 11) % r]
    for _ in range(m):
        x.append((x[-1] * x[-1] + b) % r)
    wire = [2 + i for i in range(m)] + [1]
    wit = [1, x[m]] + x[:m]
    n8 = 32

    def lc(terms):
        out = struct.pack("<I", len(terms))
        for w, v in terms:
            out += struct.pack("<I", w) + (v % r).to_bytes(n8, "little")
        return out

    body = b""
    for i in range(m):
        body += lc([(wire[i], 1)]) + lc([(wire[i], 1)]) + lc([(wire[i + 1], 1), (0, -b)])
    n_wires = len(wit)
    hdr = struct.pack("<I", n8) + r.to_bytes(n8, "little") + struct.pack("<IIII", n_wires, 1, 0, m) + struct.pack("<Q", n_wires) + struct.pack("<I", m)
    labels = b"".join(struct.pack("<Q", i) for i in range(n_wires))
    return orc.write_binfile("r1cs", 1, [(1, hdr), (2, body), (3, labels)]), wit


def prepared_ptau(curve: int, domain_size: int, tau: int, alpha: int, beta: int) -> bytes:
    """Sections 1, 4, 5, 6, 12, 13, 14, 15 of a prepared ptau, filled only where zkey_new reads: the Lagrange bases of size n
    (offset n - 1) and, for section 12, of size 2n (offset 2n - 1).  L_j(tau) = (tau^n - 1) w^j / (n (tau - w^j))."""
    ci = orc.CURVES[curve]
    r, n = ci.r, domain_size
    power = n.bit_length() - 1
    sG1, sG2 = 2 * ci.n8q, 4 * ci.n8q
    g1 = orc.g_from_affine(ci.id, 1, ci.g1_affine_bytes(ci.g1))
    g2 = orc.g_from_affine(ci.id, 2, ci.g2_affine_bytes(ci.g2))

    def lagrange(size):
        w = ci.fr_from_mont(orc.fr_root(ci.id, size.bit_length() - 1))
        zt = (pow(tau, size, r) - 1) % r
        inv_n = pow(size, -1, r)
        out, wj = [], 1
        for _ in range(size):
            out.append(zt * wj % r * inv_n % r * pow((tau - wj) % r, -1, r) % r)
            wj = wj * w % r
        return out

    def points(group, scalars):
        g = g1 if group == 1 else g2
        jac = b"".join(orc.g_times(ci.id, group, g, (s % r).to_bytes(32, "little")) for s in scalars)
        return bytes(orc.batch_to_affine(ci.id, group, jac))

    Ln, L2n = lagrange(n), lagrange(2 * n)
    sec12 = bytes((n - 1) * sG1) + points(1, Ln) + points(1, L2n)
    sec13 = bytes((n - 1) * sG2) + points(2, Ln)
    sec14 = bytes((n - 1) * sG1) + points(1, [alpha * x % r for x in Ln])
    sec15 = bytes((n - 1) * sG1) + points(1, [beta * x % r for x in Ln])
    hdr = struct.pack("<I", ci.n8q) + ci.q.to_bytes(ci.n8q, "little") + struct.pack("<II", power + 1, power + 1)
    return orc.write_binfile("ptau", 1, [(1, hdr), (4, points(1, [alpha])), (5, points(1, [beta])), (6, points(2, [beta])),
                                         (12, sec12), (13, sec13), (14, sec14), (15, sec15)])
