// Host backend for snarkjs_b200/csrc/plonk_flow.h: the PLONK control flow and the plonk.cuh element functions
// compiled with g++, bulk NTT / MSM borrowed from the CPU oracle (dlopen).  Built as a shared library and driven by
// tests/test_host_plonk.py, which compares the proof bytes with oracle/plonk.py.  Test infrastructure only.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <dlfcn.h>
#include <vector>
#include <string>
#include "../../snarkjs_b200/csrc/plonk_flow.h"
using namespace sb;

typedef int (*or_fft_t)(int, const uint8_t*, uint64_t, int, uint8_t*);
typedef int (*or_msm_t)(int, int, const uint8_t*, const uint8_t*, int, uint64_t, int, uint8_t*);
typedef int (*or_gop_t)(int, int, int, const uint8_t*, const uint8_t*, uint8_t*);
typedef int (*or_root_t)(int, int, uint8_t*);

template <class F> struct HostBackend {
    or_fft_t fft; or_msm_t msm; or_gop_t gop;
    int curve; uint32_t n8q;
    const uint8_t* ptau;
    std::vector<std::vector<F>> pow_store;
    std::string err;

    void upload(F* dst, const F* host, size_t n) { memcpy(dst, host, n * sizeof(F)); }
    void download(F* host, const F* src, size_t n) { memcpy(host, src, n * sizeof(F)); }
    void zero(F* p, size_t n) { memset(p, 0, n * sizeof(F)); }
    void copy(F* dst, const F* src, size_t n) { memmove(dst, src, n * sizeof(F)); }
    F* ntt(F* a, F* b, uint64_t n, bool inverse) { fft(curve, (const uint8_t*)a, n, inverse ? 1 : 0, (uint8_t*)b); return b; }
    int commit_plain(const F* scal, uint64_t len, uint8_t* affine) {
        std::vector<uint8_t> jac(3 * n8q), aff(3 * n8q);
        if (msm(curve, 1, ptau, (const uint8_t*)scal, 32, len, 4, jac.data())) return -9;
        if (gop(curve, 1, 2, jac.data(), nullptr, aff.data())) return -9;
        memcpy(affine, aff.data(), 2 * n8q);
        return 0;
    }
    int commit(const F* coef, uint64_t len, uint8_t* affine) {
        std::vector<F> s(len);
        for (uint64_t i = 0; i < len; i++) s[i] = F::from_mont(coef[i]);
        return commit_plain(s.data(), len, affine);
    }
    void additions(const PlonkKeyView<F>& k, F* W) {
        uint32_t lo = 0;
        for (uint32_t hi : k.level_end) { for (uint32_t j = lo; j < hi; j++) pl_addition<F>(k.add_order[j], k.add_sig, k.add_fac, W, k.nVars - k.nAdditions, k.nVars); lo = hi; }
    }
    void wires(const PlonkKeyView<F>& k, const F* W, F* A, F* B, F* C) {
        F* out[3] = {A, B, C};
        for (int j = 0; j < 3; j++) for (uint64_t i = 0; i < k.n; i++) pl_wire<F>(i, k.map[j], W, k.nVars, k.nConstraints, out[j]);
    }
    void blind(F* p, uint64_t n, const F* bf, int cnt) { pl_blind<F>(p, n, bf, cnt); }
    int z(const PlonkKeyView<F>& k, const PlonkRound<F>& r, PlonkWork<F>& w) {
        const uint64_t n = k.n;
        for (uint64_t i = 0; i < n; i++) pl_z_terms<F>(i, w.bufA, w.bufB, w.bufC, k.s_ev[0], k.s_ev[1], k.s_ev[2], k.wpow, r, w.num, w.den);
        for (uint64_t lo = 0; lo < n; lo += 16) pl_ratio_chunk<F>(w.den, w.num, w.ratio, lo, lo + 16 < n ? lo + 16 : n);
        F acc = F::one();
        for (uint64_t i = 0; i < n; i++) { w.bufZ[i] = acc; acc = F::mul(acc, w.ratio[i]); }     // exclusive product scan
        return (F::mul(w.bufZ[n - 1], w.ratio[n - 1]) == F::one()) ? 0 : 4;
    }
    void t(const PlonkKeyView<F>& k, const PlonkRound<F>& r, PlonkWork<F>& w) {
        PlonkTIn in;
        in.A = w.evA; in.B = w.evB; in.C = w.evC; in.Z = w.evZ;
        in.QM = k.q_ev[0]; in.QL = k.q_ev[1]; in.QR = k.q_ev[2]; in.QO = k.q_ev[3]; in.QC = k.q_ev[4];
        in.S1 = k.s_ev[0]; in.S2 = k.s_ev[1]; in.S3 = k.s_ev[2]; in.LAG = k.lag; in.pubA = w.bufA; in.n_public = k.nPublic;
        for (uint64_t i = 0; i < 4ull * k.n; i++) pl_t_eval<F>(i, 4ull * k.n, in, k.w4pow, r, w.T, w.Tz);
    }
    int divzh(uint64_t n, const F* t, const F* tz, F* out) { int f = 0; for (uint64_t i = 0; i < n; i++) f |= pl_divzh<F>(i, n, t, tz, out); return f; }
    void tsplit(uint64_t n, const F* t, const F& b10, const F& b11, F* T1, F* T2, F* T3) { for (uint64_t i = 0; i < n + 6; i++) pl_tsplit<F>(i, n, t, b10, b11, T1, T2, T3); }
    void make_pow(const F& base, uint64_t count, PlonkPow<F>& out, int) {
        int h = plonk_pow_h(count);
        std::vector<F> lo, hi; plonk_pow_tables<F>(base, h, (count >> h) + 1, lo, hi);
        pow_store.push_back(lo); out.lo = pow_store.back().data();
        pow_store.push_back(hi); out.hi = pow_store.back().data(); out.h = h;
    }
    F eval(const F* f, uint64_t len, const PlonkPow<F>& pw, F*, F*) {
        F s = F::zero();
        for (uint64_t i = 0; i < len; i++) s = F::add(s, F::mul(f[i], pl_pow(pw, i)));
        return s;
    }
    int quotient(const F* f, const PlonkLinIn* lin, const PlonkLin<F>* L, uint64_t n, uint64_t len, uint64_t m, const F& sub0,
                 const PlonkPow<F>& pw, const PlonkPow<F>& ipw, F* g, F* P, F* q_plain) {
        for (uint64_t i = 0; i < m; i++) {
            F x;
            if (lin) x = pl_wxi_coef<F>(i, n, *lin, *L);
            else { x = i < len ? f[i] : F::zero(); if (i == 0) x = F::sub(x, sub0); }
            g[i] = F::mul(x, pl_pow(pw, i));
        }
        F acc = F::zero();
        for (uint64_t i = 0; i < m; i++) { acc = F::add(acc, g[i]); P[i] = acc; }                    // inclusive sum scan
        for (uint64_t j = 0; j < m; j++) q_plain[j] = F::from_mont(pl_quot_coef<F>(j, m, P, ipw));
        return P[m - 1].is_zero() ? 0 : 1;
    }
};

template <class PQ, class PR>
static int prove_impl(void* so, int curve, const PlonkZkey& z, const uint8_t* witness, uint64_t n_wit, const uint8_t* blinders, uint8_t* proof, std::string& err) {
    typedef Fp<PR> F;
    HostBackend<F> be;
    be.fft = (or_fft_t)dlsym(so, "or_fr_fft"); be.msm = (or_msm_t)dlsym(so, "or_multiexp_affine"); be.gop = (or_gop_t)dlsym(so, "or_group_op");
    or_root_t root = (or_root_t)dlsym(so, "or_fr_root");
    if (!be.fft || !be.msm || !be.gop || !root) { err = "oracle symbols missing"; return -1; }
    be.curve = curve; be.n8q = z.n8q; be.ptau = z.sec[14].p;
    be.pow_store.reserve(64);
    PlonkKeyView<F> k;
    k.nVars = z.nVars; k.nPublic = z.nPublic; k.n = z.n; k.nAdditions = z.nAdditions; k.nConstraints = z.nConstraints; k.power = z.power;
    memcpy(&k.k1, z.k1, 32); memcpy(&k.k2, z.k2, 32);
    F w2; root(curve, z.power, (uint8_t*)&k.wn); root(curve, z.power + 2, (uint8_t*)&k.w4n); root(curve, 2, (uint8_t*)&w2);
    plonk_mulz_tables<F>(w2, k.z1, k.z2, k.z3);
    k.hdr_pts = z.hdr_pts; k.aff_bytes = 2 * z.n8q;
    const uint64_t n = z.n;
    // additions: split (s1, s2, f1, f2) records into the two arrays the kernels read
    std::vector<uint32_t> sig(2 * (size_t)z.nAdditions + 2), order; std::vector<F> fac(2 * (size_t)z.nAdditions + 2);
    for (uint32_t i = 0; i < z.nAdditions; i++) { memcpy(&sig[2 * i], z.sec[3].p + 72 * (size_t)i, 8); memcpy(&fac[2 * i], z.sec[3].p + 72 * (size_t)i + 8, 64); }
    plonk_addition_levels(sig.data(), z.nAdditions, z.nVars - z.nAdditions, order, k.level_end);
    k.add_sig = sig.data(); k.add_fac = fac.data(); k.add_order = order.data();
    std::vector<uint32_t> maps[3];
    for (int j = 0; j < 3; j++) { maps[j].resize(z.nConstraints + 1); memcpy(maps[j].data(), z.sec[4 + j].p, 4 * (size_t)z.nConstraints); k.map[j] = maps[j].data(); }
    std::vector<F> qc[5], qe[5], sc[3], se[3], lag;
    for (int j = 0; j < 5; j++) { qc[j].resize(n); qe[j].resize(4 * n); memcpy(qc[j].data(), z.sec[7 + j].p, 32 * n); memcpy(qe[j].data(), z.sec[7 + j].p + 32 * n, 128 * n); k.q_coef[j] = qc[j].data(); k.q_ev[j] = qe[j].data(); }
    for (int j = 0; j < 3; j++) { sc[j].resize(n); se[j].resize(4 * n); memcpy(sc[j].data(), z.sec[12].p + 160 * n * j, 32 * n); memcpy(se[j].data(), z.sec[12].p + 160 * n * j + 32 * n, 128 * n); k.s_coef[j] = sc[j].data(); k.s_ev[j] = se[j].data(); }
    const uint32_t nl = z.nPublic > 1 ? z.nPublic : 1;
    const uint32_t nl_have = (uint32_t)(z.sec[13].len / (160 * n));
    lag.assign((size_t)nl * 4 * n, F::zero());
    for (uint32_t j = 0; j < nl && j < nl_have; j++) memcpy(lag.data() + (size_t)j * 4 * n, z.sec[13].p + 160 * n * j + 32 * n, 128 * n);
    k.lag = lag.data();
    be.make_pow(k.wn, n, k.wpow, 0);
    be.make_pow(k.w4n, 4 * n, k.w4pow, 0);
    PlonkWork<F> w;
    std::vector<std::vector<F>> store;
    auto alloc = [&](size_t cnt) { store.emplace_back(cnt, F::zero()); return store.back().data(); };
    store.reserve(64);
    w.W = alloc(z.nVars + 2);
    w.bufA = alloc(n); w.bufB = alloc(n); w.bufC = alloc(n); w.bufZ = alloc(n); w.num = alloc(n); w.den = alloc(n); w.ratio = alloc(n); w.sn = alloc(n);
    w.cA = alloc(n + 8); w.cB = alloc(n + 8); w.cC = alloc(n + 8); w.cZ = alloc(n + 8); w.T1 = alloc(n + 8); w.T2 = alloc(n + 8); w.T3 = alloc(n + 8);
    w.g = alloc(n + 8); w.P = alloc(n + 8); w.scal = alloc(n + 8);
    w.evA = alloc(4 * n); w.evB = alloc(4 * n); w.evC = alloc(4 * n); w.evZ = alloc(4 * n); w.T = alloc(4 * n); w.Tz = alloc(4 * n); w.s4a = alloc(4 * n); w.s4b = alloc(4 * n);
    return plonk_prove_flow<PQ, PR>(be, k, w, witness, n_wit, blinders, proof, err);
}

extern "C" {
int hp_keccak256(const uint8_t* data, uint64_t len, uint8_t* out) { keccak256(data, len, out); return 0; }

// proof = 9 affine points (Montgomery) | 6 evaluations (Montgomery); returns 0 or a negative code with err filled
int hp_plonk_prove(const char* oracle_so, const uint8_t* zkey, uint64_t zlen, const uint8_t* witness, uint64_t n_wit, const uint8_t* blinders,
                   uint8_t* proof, char* errbuf, int errlen) {
    std::string err;
    void* so = dlopen(oracle_so, RTLD_NOW);
    if (!so) { snprintf(errbuf, errlen, "dlopen failed: %s", dlerror()); return -1; }
    PlonkZkey z;
    int rc = plonk_parse_zkey(zkey, zlen, z, err);
    if (!rc) {
        if (z.n8q == 32) rc = prove_impl<BnFq, BnFr>(so, 0, z, witness, n_wit, blinders, proof, err);
        else rc = prove_impl<BlsFq, BlsFr>(so, 1, z, witness, n_wit, blinders, proof, err);
    }
    snprintf(errbuf, errlen, "%s", err.c_str());
    return rc;
}
}
