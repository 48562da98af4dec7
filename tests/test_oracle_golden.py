"""Pins the CPU oracle against reference-produced bytes (tests/golden/*.npz)
and against independent Python-int arithmetic.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O

BN = O.BN254


def test_field_mul_vs_python_ints():
    rng = np.random.default_rng(1)
    for fid, p in ((O.F_BN_FQ, O.P_BN_Q), (O.F_BN_FR, O.P_BN_R), (O.F_BLS_FQ, O.P_BLS_Q), (O.F_BLS_FR, O.P_BLS_R)):
        n8 = 48 if fid == O.F_BLS_FQ else 32
        R = 1 << (8 * n8)
        Rinv = pow(R, -1, p)
        for _ in range(2000):
            a = int.from_bytes(rng.bytes(n8), "little") % p
            b = int.from_bytes(rng.bytes(n8), "little") % p
            ab, bb = a.to_bytes(n8, "little"), b.to_bytes(n8, "little")
            assert int.from_bytes(O.field_op(fid, 2, ab, bb), "little") == a * b * Rinv % p
            assert int.from_bytes(O.field_op(fid, 0, ab, bb), "little") == (a + b) % p
            assert int.from_bytes(O.field_op(fid, 1, ab, bb), "little") == (a - b) % p
        # edge values
        for a in (0, 1, p - 1):
            for b in (0, 1, p - 1):
                ab, bb = a.to_bytes(n8, "little"), b.to_bytes(n8, "little")
                assert int.from_bytes(O.field_op(fid, 2, ab, bb), "little") == a * b * Rinv % p
        x = (12345 * R) % p
        assert int.from_bytes(O.field_op(fid, 4, x.to_bytes(n8, "little")), "little") == pow(12345, -1, p) * R % p


def test_roots_of_unity_match_survey():
    # BN254 w[28], BLS w[32]
    ci = O.CURVES[BN]
    assert O.fr_s(BN) == 28
    assert ci.fr_from_mont(O.fr_root(BN, 28)) == 0x2a3c09f0a58a7e8500e0a7eb8ef62abc402d111e41112ed49bd61b6e725b19f0
    assert ci.fr_from_mont(O.fr_root(BN, -2)) == 5
    cb = O.CURVES[O.BLS12_381]
    assert O.fr_s(O.BLS12_381) == 32
    assert cb.fr_from_mont(O.fr_root(O.BLS12_381, 32)) == 0x212d79e5b416b6f0fd56dc8d168d6c0c4024ff270b3e0941b788f500b912f1f
    assert cb.fr_from_mont(O.fr_root(O.BLS12_381, -2)) == 5


def test_ntt_goldens(golden):
    g = golden("ntt_goldens.npz")
    labels = sorted({k[:-5] for k in g if k.endswith("_coef")})
    assert len(labels) == 9
    for lab in labels:
        coef, evals = g[lab + "_coef"], g[lab + "_evals"]
        n = coef.size // 32
        padded = np.concatenate([coef, np.zeros(3 * n * 32, dtype=np.uint8)])
        assert np.array_equal(O.fr_fft(BN, padded), evals), lab
        back = O.fr_fft(BN, evals, inverse=True)
        assert np.array_equal(back[:n * 32], coef), lab
        assert not back[n * 32:].any(), lab


def test_ntt_vs_naive_dft():
    ci = O.CURVES[BN]
    n = 64
    rng = np.random.default_rng(2)
    xs = [int.from_bytes(rng.bytes(32), "little") % ci.r for _ in range(n)]
    buf = b"".join(ci.fr_to_mont(x) for x in xs)
    w = ci.fr_from_mont(O.fr_root(BN, 6))
    out = O.fr_fft(BN, buf).tobytes()
    for j in range(n):
        want = sum(xs[i] * pow(w, i * j, ci.r) for i in range(n)) % ci.r
        assert ci.fr_from_mont(out[32 * j:32 * j + 32]) == want


def test_g1_msm_goldens(golden):
    g = golden("msm_g1_goldens.npz")
    names = sorted(k[:-7] for k in g if k.endswith("_commit"))
    assert len(names) == 12
    for nm in names:
        tag = nm.split("_")[0]
        scal = O.batch_convert(O.F_BN_FR, False, g[nm + "_coef_mont"])
        n = scal.size // 32
        jac = O.multiexp_affine(BN, 1, g[tag + "_ptau"][:64 * n], scal)
        assert O.g_to_affine(BN, 1, jac) == g[nm + "_commit"].tobytes(), nm


def _lagrange_scalars(k, j):
    ci = O.CURVES[BN]
    n = 1 << k
    w = ci.fr_from_mont(O.fr_root(BN, k))
    winv = pow(w, -1, ci.r)
    ninv = pow(n, -1, ci.r)
    return b"".join((pow(winv, i * j, ci.r) * ninv % ci.r).to_bytes(32, "little") for i in range(n))


def test_ptau_lagrange_goldens_g1_g2(golden):
    g = golden("ptau_goldens.npz")
    for grp, key, sz in ((1, "g1", 64), (2, "g2", 128)):
        base = g["tauG1"] if grp == 1 else g["tauG2"]
        for idx, (k, j) in enumerate(g[key + "_picks"]):
            n = 1 << int(k)
            jac = O.multiexp_affine(BN, grp, base[:sz * n], _lagrange_scalars(int(k), int(j)))
            assert O.g_to_affine(BN, grp, jac) == g[key + "_expected"][idx * sz:(idx + 1) * sz].tobytes(), (key, k, j)


def test_msm_pippenger_vs_naive_small():
    rng = np.random.default_rng(3)
    for grp in (1, 2):
        for curve in (BN, O.BLS12_381):
            n = 37
            bases = O.gen_points(curve, grp, 7, n)
            sc = O.random_scalars(11, n, O.CURVES[curve].r)
            a = O.multiexp_affine(curve, grp, bases, sc, 3)
            b = O.multiexp_naive(curve, grp, bases, sc)
            assert O.g_eq(curve, grp, a, b)


def test_error_conventions():
    with pytest.raises(ValueError, match="fft must be multiple of 2"):
        O.fr_fft(BN, bytes(32 * 3))
    with pytest.raises(ValueError, match="Scalar size does not match"):
        O.multiexp_affine(BN, 1, bytes(64 * 3), bytes(32 * 3 + 1))
    assert O.multiexp_affine(BN, 1, b"", b"") == O.group_zero(BN, 1)


def test_groth16_oracle_proof_verifies(golden):
    """Mirrors test/fullprocess.js:120-133: prove -> verify == true; aliased public input rejected."""
    g = golden("groth16_case.npz")
    zkey, wt = g["zkey"].tobytes(), g["wtns"].tobytes()
    ci = O.CURVES[BN]
    r, s = ci.fr_to_mont(123456789), ci.fr_to_mont(987654321)
    proof, pub = O.groth16_prove(zkey, wt, r, s)
    assert pub[1] == 11 and len(pub) == 2
    vk = O.zkey_vk(zkey)
    assert O.groth16_verify(vk, pub, proof)
    assert not O.groth16_verify(vk, [pub[0] + ci.r] + pub[1:], proof)        # aliased public signal
    assert not O.groth16_verify(vk, [(pub[0] + 1) % ci.r] + pub[1:], proof)
    # different (r,s) -> different but still valid proof
    proof2, _ = O.groth16_prove(zkey, wt, ci.fr_to_mont(5), ci.fr_to_mont(7))
    assert proof2 != proof and O.groth16_verify(vk, pub, proof2)
