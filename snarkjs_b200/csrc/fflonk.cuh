// fflonk.cuh — the per-element work of snarkjs' fflonk prover (src/fflonk_prove.js) between its NTTs and MSMs, in the
// style of plonk.cuh: every loop body of the reference is an SB_HD function of the element index (one CUDA thread per
// index on the device, a plain loop in tests/host/host_fflonk.cpp), and its recurrences become scans:
//
//   f / (X^m - b)   (Polynomial.divByZerofier, polynomial.js:617-660) — with f(X) = sum_j X^j f_j(X^m) this is m independent
//                   divisions f_j(Y) / (Y - b) on the stride-m subsequences: g = f[k] b^(k div m) stored class-major,
//                   a segmented inclusive sum scan P, q[k] = (P_last(class) - P[k]) b^-(k div m + 1); the class totals
//                   are the remainder and must vanish.
//   f / (X - y)     (Polynomial.divBy, :341-360) — the m = 1 case, shared with plonk.cuh (pl_quot_coef).
#pragma once
#include "plonk.cuh"

namespace sb {

// ---------------------------------------------------------------------------------------------- round 1
// computeT0 (fflonk_prove.js:415-504): T0 * Z_H on the 4n domain.  PlonkTIn carries A B C QM QL QR QO QC LAG pubA n_public.
template <class F> SB_HD void ff_t0(uint64_t i, uint64_t n4, const PlonkTIn& in, F* T0) {
    const F a = pl_ld((const F*)in.A + i), b = pl_ld((const F*)in.B + i), c = pl_ld((const F*)in.C + i);
    F pi = F::zero();
    for (uint32_t j = 0; j < in.n_public; j++)
        pi = F::sub(pi, F::mul(pl_ld((const F*)in.LAG + (uint64_t)j * n4 + i), pl_ld((const F*)in.pubA + j)));
    F t = F::mul(a, pl_ld((const F*)in.QL + i));
    t = F::add(t, F::mul(b, pl_ld((const F*)in.QR + i)));
    t = F::add(t, F::mul(F::mul(a, b), pl_ld((const F*)in.QM + i)));
    t = F::add(t, F::mul(c, pl_ld((const F*)in.QO + i)));
    t = F::add(t, F::add(pl_ld((const F*)in.QC + i), pi));
    pl_st(T0 + i, t);
}
// wire blinding (:375-380): the reference writes the blinders' Montgomery bytes into the *plain* evaluation buffer before
// batchToMontgomery, so the evaluation is toMontgomery(raw bytes)
template <class F> SB_HD void ff_wire_blind(F* buf, uint64_t n, const F& b_lo_raw, const F& b_hi_raw) {
    pl_st(buf + n - 2, F::to_mont(b_lo_raw));
    pl_st(buf + n - 1, F::to_mont(b_hi_raw));
}

// ---------------------------------------------------------------------------------------------- round 2
// computeT1 (:667-718) on the 2n domain: (z - 1) L1 and its blinding part.  evZ and lag1 are 4n-point arrays (stride 2).
template <class F> SB_HD void ff_t1(uint64_t i, const F* evZ, const F* lag1, const PlonkPow<F>& w2pow, const PlonkRound<F>& r, F* T1, F* T1z) {
    const F om = pl_pow(w2pow, i);
    const F zp = F::add(F::mul(F::add(F::mul(r.b[7], om), r.b[8]), om), r.b[9]);
    const F l1 = pl_ld(lag1 + 2 * i);
    pl_st(T1 + i, F::mul(F::sub(pl_ld(evZ + 2 * i), F::one()), l1));
    pl_st(T1z + i, F::mul(zp, l1));
}
// computeT2 (:720-815) on the 4n domain.  PlonkTIn carries A B C Z S1 S2 S3.
template <class F> SB_HD void ff_t2(uint64_t i, uint64_t n4, const PlonkTIn& in, const PlonkPow<F>& w4pow, const PlonkRound<F>& r, F* T2, F* T2z) {
    const F a = pl_ld((const F*)in.A + i), b = pl_ld((const F*)in.B + i), c = pl_ld((const F*)in.C + i);
    const F z = pl_ld((const F*)in.Z + i), zw = pl_ld((const F*)in.Z + ((i + 4) & (n4 - 1)));
    const F om = pl_pow(w4pow, i), omW = F::mul(om, r.wn);
    const F zp = F::add(F::mul(F::add(F::mul(r.b[7], om), r.b[8]), om), r.b[9]);
    const F zWp = F::add(F::mul(F::add(F::mul(r.b[7], omW), r.b[8]), omW), r.b[9]);
    const F betaX = F::mul(r.beta, om);
    F e1 = F::mul(F::add(F::add(a, betaX), r.gamma), F::add(F::add(b, F::mul(betaX, r.k1)), r.gamma));
    e1 = F::mul(e1, F::add(F::add(c, F::mul(betaX, r.k2)), r.gamma));
    F e2 = F::mul(F::add(F::add(a, F::mul(r.beta, pl_ld((const F*)in.S1 + i))), r.gamma), F::add(F::add(b, F::mul(r.beta, pl_ld((const F*)in.S2 + i))), r.gamma));
    e2 = F::mul(e2, F::add(F::add(c, F::mul(r.beta, pl_ld((const F*)in.S3 + i))), r.gamma));
    pl_st(T2 + i, F::sub(F::mul(e1, z), F::mul(e2, zw)));
    pl_st(T2z + i, F::sub(F::mul(e1, zp), F::mul(e2, zWp)));
}
// divByZerofier(n, 1) (polynomial.js:617-660) on a polynomial of `blocks` * n coefficients, then add the blinding part tz
// (may be null) and check the degree bound: coefficients at index >= bound must be zero afterwards.
// Thread i < n owns coefficients i, n+i, 2n+i, ...  Returns 1 ("Polynomial is not divisible") or 2 (degree) or 0.
template <class F> SB_HD int ff_divzh(uint64_t i, uint64_t n, int blocks, const F* t, const F* tz, F* out, uint64_t bound) {
    int bad = 0;
    F c = F::neg(pl_ld(t + i));
    for (int k = 0; k < blocks; k++) {
        if (k) c = F::sub(c, pl_ld(t + (uint64_t)k * n + i));
        if (k == blocks - 1 && !c.is_zero()) bad = 1;                 // the top n coefficients must vanish
        F o = tz ? F::add(c, pl_ld(tz + (uint64_t)k * n + i)) : c;
        if ((uint64_t)k * n + i >= bound && !o.is_zero() && !bad) bad = 2;
        pl_st(out + (uint64_t)k * n + i, o);
    }
    return bad;
}
// CPolynomial.getPolynomial (cpolynomial.js:52-72): out[i m + j] = p_j[i]
struct FfParts { const void* p[4]; uint64_t len[4]; int m; };
template <class F> SB_HD void ff_interleave(uint64_t k, const FfParts& parts, F* out) {
    const uint64_t j = k % parts.m, i = k / parts.m;
    pl_st(out + k, i < parts.len[j] ? pl_ld((const F*)parts.p[j] + i) : F::zero());
}

// ---------------------------------------------------------------------------------------------- round 4
template <class F> struct FfSmall { F c[8]; int len; };               // R0 / R1 / R2: at most 8 coefficients
// numerator coefficient k of (f - R) * scale, times b^(k div m), stored class-major: G[(k mod m) rows + k div m]
template <class F> SB_HD void ff_qm_g(uint64_t k, const F* f, uint64_t len, const FfSmall<F>& R, const F& scale, int m, uint64_t rows,
                                      const PlonkPow<F>& bpow, F* G) {
    F x = k < len ? pl_ld(f + k) : F::zero();
    if (k < (uint64_t)R.len) x = F::sub(x, R.c[k]);
    x = F::mul(x, scale);
    const uint64_t j = k % m, t = k / m;
    pl_st(G + j * rows + t, F::mul(x, pl_pow(bpow, t)));
}
// quotient coefficient k from the segmented inclusive sums P (class-major); the top row is zero.
// Returns nonzero when class k mod m leaves a remainder (checked once per class, by the thread of its first element).
template <class F> SB_HD int ff_qm_q(uint64_t k, int m, uint64_t rows, const F* P, const PlonkPow<F>& ibpow, F* q) {
    const uint64_t j = k % m, t = k / m;
    const F total = pl_ld(P + j * rows + rows - 1);
    int bad = (t == 0 && !total.is_zero()) ? 1 : 0;
    pl_st(q + k, t + 1 >= rows ? F::zero() : F::mul(F::sub(total, pl_ld(P + j * rows + t)), pl_pow(ibpow, t + 1)));
    return bad;
}

// ---------------------------------------------------------------------------------------------- round 5
template <class F> struct FfLin { F pre0, pre1, pre2, r0y, r1y, r2y, zty, zts2y_inv; };
// computeL (:1101-1162) then mulScalar(1 / ZTS2(y)) (:1077-1079): coefficient k of
//   [preL0 (C0 - R0(y)) + preL1 (C1 - R1(y)) + preL2 (C2 - R2(y)) - ZT(y) F] / ZTS2(y)
template <class F> SB_HD F ff_l_coef(uint64_t k, const F* C0, uint64_t l0, const F* C1, uint64_t l1, const F* C2, uint64_t l2, const F* Fp, uint64_t lf,
                                     const FfLin<F>& L) {
    F c0 = pl_at<F>(C0, k, l0), c1 = pl_at<F>(C1, k, l1), c2 = pl_at<F>(C2, k, l2);
    if (k == 0) { c0 = F::sub(c0, L.r0y); c1 = F::sub(c1, L.r1y); c2 = F::sub(c2, L.r2y); }
    F x = F::mul(c0, L.pre0);
    x = F::add(x, F::mul(c1, L.pre1));
    x = F::add(x, F::mul(c2, L.pre2));
    x = F::sub(x, F::mul(pl_at<F>(Fp, k, lf), L.zty));
    return F::mul(x, L.zts2y_inv);
}

#ifdef __CUDACC__
template <class F> __global__ void __launch_bounds__(128) k_ff_t0(uint64_t n4, PlonkTIn in, F* T0) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n4) ff_t0<F>(i, n4, in, T0);
}
template <class F> struct FfBlind6 { F raw[6]; };
template <class F> __global__ void k_ff_wire_blind(F* A, F* B, F* C, uint64_t n, FfBlind6<F> b) {
    if (blockIdx.x == 0 && threadIdx.x < 3) { F* bufs[3] = {A, B, C}; ff_wire_blind<F>(bufs[threadIdx.x], n, b.raw[2 * threadIdx.x], b.raw[2 * threadIdx.x + 1]); }
}
template <class F> __global__ void k_ff_t1(uint64_t n2, const F* evZ, const F* lag1, PlonkPow<F> w2pow, PlonkRound<F> r, F* T1, F* T1z) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n2) ff_t1<F>(i, evZ, lag1, w2pow, r, T1, T1z);
}
template <class F> __global__ void __launch_bounds__(128) k_ff_t2(uint64_t n4, PlonkTIn in, PlonkPow<F> w4pow, PlonkRound<F> r, F* T2, F* T2z) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n4) ff_t2<F>(i, n4, in, w4pow, r, T2, T2z);
}
template <class F> __global__ void k_ff_divzh(uint64_t n, int blocks, const F* t, const F* tz, F* out, uint64_t bound, int* flag) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) { int bad = ff_divzh<F>(i, n, blocks, t, tz, out, bound); if (bad) atomicOr(flag, bad); }
}
template <class F> __global__ void k_ff_interleave(uint64_t total, FfParts parts, F* out) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k < total) ff_interleave<F>(k, parts, out);
}
template <class F> struct FfOne { F x; };
template <class F> __global__ void k_ff_qm_g(uint64_t total, const F* f, uint64_t len, FfSmall<F> R, FfOne<F> scale, int m, uint64_t rows, PlonkPow<F> bpow, F* G) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k < total) ff_qm_g<F>(k, f, len, R, scale.x, m, rows, bpow, G);
}
template <class F> __global__ void k_ff_qm_q(uint64_t total, int m, uint64_t rows, const F* P, PlonkPow<F> ibpow, F* q, int* flag) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k < total && ff_qm_q<F>(k, m, rows, P, ibpow, q)) atomicOr(flag, 1);
}
template <class F> __global__ void k_ff_add3(uint64_t total, const F* a, const F* b, const F* c, F* out) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k < total) pl_st(out + k, F::add(F::add(pl_ld(a + k), pl_ld(b + k)), pl_ld(c + k)));
}
template <class F> __global__ void k_ff_l(uint64_t total, const F* C0, uint64_t l0, const F* C1, uint64_t l1, const F* C2, uint64_t l2, const F* Fp, uint64_t lf,
                                          FfLin<F> L, PlonkPow<F> ypow, F* g) {
    uint64_t k = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (k < total) pl_st(g + k, F::mul(ff_l_coef<F>(k, C0, l0, C1, l1, C2, l2, Fp, lf, L), pl_pow(ypow, k)));
}
#endif  // __CUDACC__

}  // namespace sb
