// Host backend for snarkjs_b200/csrc/plonk_flow.h: the PLONK control flow and the plonk.cuh element functions
// compiled with g++, bulk NTT / MSM borrowed from the CPU oracle (dlopen).  Built as a shared library and driven by
// tests/test_host_plonk.py, which compares the proof bytes with oracle/plonk.py.  Test infrastructure only.
#include "host_backend.h"

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
