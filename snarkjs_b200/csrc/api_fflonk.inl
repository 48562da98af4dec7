// api_fflonk.inl — fflonk proving on the device (part of api.cu's translation unit, after api_plonk.inl).
//
//   sb_fflonk_load    zkey sections 3-17 -> HBM (selectors / sigmas / Lagrange in coefficient and 4n-evaluation form, wire
//                     maps, additions by dependency level, C0's 8n coefficients, the 9n + 18 PTau bases with window tables)
//   sb_fflonk_prove   fflonk_flow.h's five rounds (src/fflonk_prove.js:51-1286) on the CUDA backend below: the PLONK
//                     backend's shared steps (additions, wires, computeZ, evaluations, commitments) plus the kernels of
//                     fflonk.cuh and a segmented cub scan for the divisions by X^m - b.
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include "fflonk_flow.h"

namespace {

struct FflonkKeyDev {
    FflonkZkey z;
    std::vector<uint8_t> hdr;
    std::vector<void*> allocs;
    uint32_t* d_add_sig = nullptr; void* d_add_fac = nullptr; uint32_t* d_add_order = nullptr; std::vector<uint32_t> level_end;
    uint32_t* d_map[3] = {nullptr, nullptr, nullptr};
    void* d_q_coef[5] = {nullptr}; void* d_q_ev[5] = {nullptr}; void* d_s_coef[3] = {nullptr}; void* d_s_ev[3] = {nullptr}; void* d_lag = nullptr; void* d_c0 = nullptr;
    void* d_ptau = nullptr; void* t_ptau = nullptr; MsmGeom gp{};
    void* d_wtab[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}}; int wtab_h[3] = {0, 0, 0};   // powers of w_n, w_2n, w_4n
    void* d_pow[6][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    uint64_t pow_nhi = 0; int pow_h = 0;
    void* work[32] = {nullptr};                  // order of FflonkWork
    void* d_cub = nullptr; size_t cub_bytes = 0;
    int* d_flag = nullptr; void* d_red = nullptr;
    bool c0_is_interleave = false;
    std::vector<uint8_t> host_pub; uint64_t n_wit_resident = 0;   // as in PlonkKeyDev
    void* commit_scratch() const { return work[31]; }   // FflonkWork::scal
};

void fflonk_free_key(FflonkKeyDev* k) {
    for (void* p : k->allocs) if (p) cudaFree(p);
    if (k->t_ptau) cudaFree(k->t_ptau);
    delete k;
}

struct FfRowKey {    // segment id of the class-major layout: element idx belongs to class idx / rows
    uint64_t rows;
    __host__ __device__ __forceinline__ uint32_t operator()(uint64_t idx) const { return (uint32_t)(idx / rows); }
};
typedef thrust::transform_iterator<FfRowKey, thrust::counting_iterator<uint64_t>> FfKeyIter;

template <class F> struct CudaFflonkBackend : CudaPlonkBackend<F, FflonkKeyDev> {
    typedef CudaPlonkBackend<F, FflonkKeyDev> Base;
    using Base::c; using Base::key; using Base::st; using Base::launched; using Base::note; using Base::grid; using Base::read_flag;

    void wire_blind(F* A, F* B, F* C, uint64_t n, const F raw[6]) {
        FfBlind6<F> b; for (int i = 0; i < 6; i++) b.raw[i] = raw[i];
        k_ff_wire_blind<F><<<1, 32, 0, st()>>>(A, B, C, n, b); launched("k_ff_wire_blind");
    }
    void t0(const PlonkTIn& in, uint64_t n4, F* T0) { k_ff_t0<F><<<grid(n4, 128), 128, 0, st()>>>(n4, in, T0); launched("k_ff_t0"); }
    void t1(uint64_t n2, const F* evZ, const F* lag1, const PlonkPow<F>& w2pow, const PlonkRound<F>& r, F* T1, F* T1z) {
        k_ff_t1<F><<<grid(n2, 128), 128, 0, st()>>>(n2, evZ, lag1, w2pow, r, T1, T1z); launched("k_ff_t1");
    }
    void t2(const PlonkTIn& in, uint64_t n4, const PlonkPow<F>& w4pow, const PlonkRound<F>& r, F* T2, F* T2z) {
        k_ff_t2<F><<<grid(n4, 128), 128, 0, st()>>>(n4, in, w4pow, r, T2, T2z); launched("k_ff_t2");
    }
    int divzh_n(uint64_t n, int blocks, const F* t, const F* tz, F* out, uint64_t bound) {
        k_ff_divzh<F><<<grid(n, 128), 128, 0, st()>>>(n, blocks, t, tz, out, bound, key->d_flag); launched("k_ff_divzh");
        return read_flag();
    }
    void interleave(const FfParts& parts, uint64_t total, F* out) { k_ff_interleave<F><<<grid(total, 256), 256, 0, st()>>>(total, parts, out); launched("k_ff_interleave"); }
    int quot_m(const F* f, uint64_t len, const FfSmall<F>& R, const F& scale, int m, uint64_t rows, const PlonkPow<F>& bpow, const PlonkPow<F>& ibpow, F* G, F* P, F* q) {
        const uint64_t total = rows * (uint64_t)m;
        FfOne<F> sc; sc.x = scale;
        k_ff_qm_g<F><<<grid(total, 256), 256, 0, st()>>>(total, f, len, R, sc, m, rows, bpow, G); launched("k_ff_qm_g");
        FfRowKey rk; rk.rows = rows;
        FfKeyIter keys(thrust::counting_iterator<uint64_t>(0), rk);
        size_t bytes = key->cub_bytes;
        note(cub::DeviceScan::InclusiveScanByKey(key->d_cub, bytes, keys, (const F*)G, P, FrAddOp(), (int)total, cub::Equality(), st()), "cub InclusiveScanByKey"); c->launches += 2;
        k_ff_qm_q<F><<<grid(total, 256), 256, 0, st()>>>(total, m, rows, P, ibpow, q, key->d_flag); launched("k_ff_qm_q");
        return read_flag();
    }
    void add3(uint64_t total, const F* a, const F* b, const F* cc, F* out) { k_ff_add3<F><<<grid(total, 256), 256, 0, st()>>>(total, a, b, cc, out); launched("k_ff_add3"); }
    int quot_l(uint64_t total, const F* C0, uint64_t l0, const F* C1, uint64_t l1, const F* C2, uint64_t l2, const F* Fp, uint64_t lf,
               const FfLin<F>& L, const PlonkPow<F>& ypow, const PlonkPow<F>& iypow, F* g, F* P, F* q_plain) {
        k_ff_l<F><<<grid(total, 128), 128, 0, st()>>>(total, C0, l0, C1, l1, C2, l2, Fp, lf, L, ypow, g); launched("k_ff_l");
        size_t bytes = key->cub_bytes;
        note(cub::DeviceScan::InclusiveScan(key->d_cub, bytes, (const F*)g, P, FrAddOp(), (int)total, st()), "cub InclusiveScan"); c->launches += 2;
        k_pl_quot<F><<<grid(total, 256), 256, 0, st()>>>(total, P, iypow, q_plain, key->d_flag); launched("k_pl_quot");
        return read_flag();
    }
};

template <class F> cudaError_t fflonk_cub_bytes(uint64_t items, size_t* out) {
    size_t a = 0, b = 0, d = 0, e2 = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveScan(nullptr, a, (const F*)nullptr, (F*)nullptr, FrMulOp(), F::one(), (int)items, (cudaStream_t)0);
    if (e != cudaSuccess) return e;
    e = cub::DeviceScan::InclusiveScan(nullptr, b, (const F*)nullptr, (F*)nullptr, FrAddOp(), (int)items, (cudaStream_t)0);
    if (e != cudaSuccess) return e;
    e = cub::DeviceReduce::Reduce(nullptr, d, (const F*)nullptr, (F*)nullptr, (int)items, FrAddOp(), F::zero(), (cudaStream_t)0);
    if (e != cudaSuccess) return e;
    FfRowKey rk; rk.rows = 1;
    FfKeyIter keys(thrust::counting_iterator<uint64_t>(0), rk);
    e = cub::DeviceScan::InclusiveScanByKey(nullptr, e2, keys, (const F*)nullptr, (F*)nullptr, FrAddOp(), (int)items, cub::Equality(), (cudaStream_t)0);
    if (e != cudaSuccess) return e;
    *out = std::max(std::max(a, b), std::max(d, e2)) + 256;
    return cudaSuccess;
}

template <class PR> int fflonk_load_impl(sb_ctx* c, const uint8_t* zkey, uint64_t zlen, uint64_t* handle) {
    typedef Fp<PR> F;
    std::string err;
    FflonkZkey z;
    if (fflonk_parse_zkey(zkey, zlen, z, err)) return fail(c, SB_ERR_FORMAT, err);
    if (!modulus_matches(z.q, z.n8q, c->curve, false) || !modulus_matches(z.r, z.n8r, c->curve, true)) return fail(c, SB_ERR_ARG, "zkey curve does not match the context curve");
    if (z.power + 2 > c->fr_s) return fail(c, SB_ERR_ARG, "domain too large for the 2-adicity of Fr");
    FflonkKeyDev* k = new FflonkKeyDev();
    k->hdr.assign(z.sec[2].p, z.sec[2].p + z.sec[2].len);
    k->z = z;
    k->c0_is_interleave = fflonk_c0_is_interleave(z);
    { const ptrdiff_t d = k->hdr.data() - z.sec[2].p;
      k->z.q += d; k->z.r += d; k->z.k1 += d; k->z.k2 += d; k->z.w3 += d; k->z.w4 += d; k->z.w8 += d; k->z.wr += d; k->z.X_2 += d; k->z.C0 += d; }
    for (auto& s : k->z.sec) s = PlonkZkey::Sec();
    const uint64_t n = z.n, sd = n * 32;
    bool ok = true;
    auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (!ok) return nullptr; if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); ok = false; return nullptr; } k->allocs.push_back(p); return p; };
    auto put = [&](void* dst, const void* src, size_t bytes) { if (ok && bytes && h2d(c, dst, src, bytes) != cudaSuccess) ok = false; };
    {
        const uint32_t na = z.nAdditions;
        std::vector<uint32_t> sig(2 * (size_t)na + 2), order; std::vector<F> fac(2 * (size_t)na + 2);
        for (uint32_t i = 0; i < na; i++) { memcpy(&sig[2 * (size_t)i], z.sec[3].p + 72 * (size_t)i, 8); memcpy(&fac[2 * (size_t)i], z.sec[3].p + 72 * (size_t)i + 8, 64); }
        plonk_addition_levels(sig.data(), na, z.nVars - na, order, k->level_end);
        k->d_add_sig = (uint32_t*)dalloc(sig.size() * 4); k->d_add_fac = dalloc(fac.size() * 32); k->d_add_order = (uint32_t*)dalloc((order.size() + 1) * 4);
        put(k->d_add_sig, sig.data(), sig.size() * 4); put(k->d_add_fac, fac.data(), fac.size() * 32); put(k->d_add_order, order.data(), order.size() * 4);
        if (ok) cudaStreamSynchronize(c->stream);
    }
    for (int j = 0; j < 3; j++) { k->d_map[j] = (uint32_t*)dalloc((size_t)z.nConstraints * 4 + 4); put(k->d_map[j], z.sec[4 + j].p, (size_t)z.nConstraints * 4); }
    for (int j = 0; j < 5; j++) { k->d_q_coef[j] = dalloc(sd); k->d_q_ev[j] = dalloc(4 * sd); put(k->d_q_coef[j], z.sec[7 + j].p, sd); put(k->d_q_ev[j], z.sec[7 + j].p + sd, 4 * sd); }
    for (int j = 0; j < 3; j++) { k->d_s_coef[j] = dalloc(sd); k->d_s_ev[j] = dalloc(4 * sd); put(k->d_s_coef[j], z.sec[12 + j].p, sd); put(k->d_s_ev[j], z.sec[12 + j].p + sd, 4 * sd); }
    {
        const uint32_t nl = z.nPublic > 1 ? z.nPublic : 1;
        k->d_lag = dalloc((size_t)nl * 4 * sd);
        for (uint32_t j = 0; j < nl; j++) put((uint8_t*)k->d_lag + (size_t)j * 4 * sd, z.sec[15].p + 5 * sd * j + sd, 4 * sd);
    }
    k->d_c0 = dalloc(8 * sd); put(k->d_c0, z.sec[17].p, 8 * sd);
    const uint64_t npts = 9 * n + 18;
    k->d_ptau = dalloc(npts * c->g1.aff_bytes); put(k->d_ptau, z.sec[16].p, npts * c->g1.aff_bytes);
    {
        std::vector<F> lo, hi;
        for (int t = 0; t < 3; t++) {     // w_n, w_2n, w_4n
            F w; memcpy(&w, c->roots[z.power + t].data(), 32);
            const uint64_t cnt = n << t;
            k->wtab_h[t] = plonk_pow_h(cnt); plonk_pow_tables<F>(w, k->wtab_h[t], (cnt >> k->wtab_h[t]) + 1, lo, hi);
            k->d_wtab[t][0] = dalloc(lo.size() * 32); k->d_wtab[t][1] = dalloc(hi.size() * 32);
            put(k->d_wtab[t][0], lo.data(), lo.size() * 32); put(k->d_wtab[t][1], hi.data(), hi.size() * 32);
            if (ok) cudaStreamSynchronize(c->stream);
        }
        const uint64_t big = 9 * n + PLONK_PAD;
        k->pow_h = plonk_pow_h(big); k->pow_nhi = (big >> k->pow_h) + 1;
        for (int s = 0; s < 6; s++) { k->d_pow[s][0] = dalloc(((size_t)1 << k->pow_h) * 32); k->d_pow[s][1] = dalloc(k->pow_nhi * 32); }
    }
    // work buffers in the order of FflonkWork
    {
        const size_t pad = PLONK_PAD * 32;
        const size_t sizes[32] = {((size_t)z.nVars + 2) * 32,
                                  sd, sd, sd, sd, sd, sd, sd,                     // bufA bufB bufC bufZ num den ratio
                                  sd, sd, sd,                                     // pA pB pC
                                  sd + pad,                                       // cZ
                                  4 * sd, 4 * sd, 4 * sd, 4 * sd, 4 * sd, 4 * sd, 4 * sd, 4 * sd,   // evA evB evC evZ T Tz s4a s4b
                                  4 * sd, 4 * sd,                                 // pT0 pT2
                                  2 * sd,                                         // pT1
                                  8 * sd,                                         // C1
                                  9 * sd + pad, 9 * sd + pad, 9 * sd + pad, 9 * sd + pad, 9 * sd + pad, 9 * sd + pad, 9 * sd + pad,   // C2 Fq F1 F2 G P scal
                                  16};
        for (int i = 0; i < 31; i++) k->work[i] = dalloc(sizes[i]);
        k->work[31] = k->work[30];                                               // commit scratch = FflonkWork::scal
    }
    k->d_flag = (int*)dalloc(16); k->d_red = dalloc(64);
    if (ok) { cudaMemsetAsync(k->d_flag, 0, 16, c->stream);
        if (fflonk_cub_bytes<F>(9 * n + PLONK_PAD, &k->cub_bytes) != cudaSuccess) ok = false; else k->d_cub = dalloc(k->cub_bytes); }
    if (ok && cudaStreamSynchronize(c->stream) != cudaSuccess) ok = false;
    if (!ok) { cudaGetLastError(); fflonk_free_key(k); return fail(c, SB_ERR_NOMEM, "fflonk key does not fit in device memory"); }
    if (want_precomp(c, npts)) {
        int rc = build_table(c, c->g1, k->d_ptau, npts, &k->t_ptau, &k->gp);
        if (rc) { fflonk_free_key(k); return rc; }
    }
    c->fflonk_keys.push_back(k);
    *handle = c->fflonk_keys.size();
    return 0;
}

template <class PQ, class PR> int fflonk_prove_impl(sb_ctx* c, FflonkKeyDev* kd, const uint8_t* witness, uint64_t n_witness, const uint8_t* blinders, uint8_t* proof) {
    typedef Fp<PR> F;
    const FflonkZkey& z = kd->z;
    FflonkKeyView<F> k;
    k.nVars = z.nVars; k.nPublic = z.nPublic; k.n = z.n; k.nAdditions = z.nAdditions; k.nConstraints = z.nConstraints; k.power = z.power;
    memcpy(&k.k1, z.k1, 32); memcpy(&k.k2, z.k2, 32); memcpy(&k.w3, z.w3, 32); memcpy(&k.w4, z.w4, 32); memcpy(&k.w8, z.w8, 32); memcpy(&k.wr, z.wr, 32);
    memcpy(&k.wn, c->roots[z.power].data(), 32);
    k.c0_point = z.C0; k.aff_bytes = c->g1.aff_bytes;
    k.add_sig = kd->d_add_sig; k.add_fac = (const F*)kd->d_add_fac; k.add_order = kd->d_add_order; k.level_end = kd->level_end;
    for (int j = 0; j < 3; j++) { k.map[j] = kd->d_map[j]; k.s_coef[j] = (const F*)kd->d_s_coef[j]; k.s_ev[j] = (const F*)kd->d_s_ev[j]; }
    for (int j = 0; j < 5; j++) { k.q_coef[j] = (const F*)kd->d_q_coef[j]; k.q_ev[j] = (const F*)kd->d_q_ev[j]; }
    k.lag = (const F*)kd->d_lag; k.c0 = (const F*)kd->d_c0; k.c0_is_interleave = kd->c0_is_interleave;
    PlonkPow<F>* tabs[3] = {&k.wpow, &k.w2pow, &k.w4pow};
    for (int t = 0; t < 3; t++) { tabs[t]->lo = (const F*)kd->d_wtab[t][0]; tabs[t]->hi = (const F*)kd->d_wtab[t][1]; tabs[t]->h = kd->wtab_h[t]; }
    FflonkWork<F> w;
    {
        int wi = 0; auto nx = [&]() { return (F*)kd->work[wi++]; };
        w.W = nx();
        w.bufA = nx(); w.bufB = nx(); w.bufC = nx(); w.bufZ = nx(); w.num = nx(); w.den = nx(); w.ratio = nx();
        w.pA = nx(); w.pB = nx(); w.pC = nx();
        w.cZ = nx();
        w.evA = nx(); w.evB = nx(); w.evC = nx(); w.evZ = nx(); w.T = nx(); w.Tz = nx(); w.s4a = nx(); w.s4b = nx();
        w.pT0 = nx(); w.pT2 = nx();
        w.pT1 = nx();
        w.C1 = nx();
        w.C2 = nx(); w.Fq = nx(); w.F1 = nx(); w.F2 = nx(); w.G = nx(); w.P = nx(); w.scal = nx();
    }
    CudaFflonkBackend<F> be; be.c = c; be.key = kd;
    std::string err;
    const size_t pub_bytes = ((size_t)z.nPublic + 1) * 32;
    if (!witness) { witness = kd->host_pub.data(); n_witness = kd->n_wit_resident; be.resident_dst = w.W; }
    else { kd->n_wit_resident = 0; if (n_witness > z.nPublic) kd->host_pub.assign(witness, witness + pub_bytes); }
    tick(c, 0);
    prof_begin(c);
    int rc = fflonk_prove_flow<PQ, PR>(be, k, w, witness, n_witness, blinders, proof, err);
    tick(c, 1);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    prof_end(c);
    if (!be.rc && rc == 0 && e == cudaSuccess) kd->n_wit_resident = n_witness;
    if (be.rc) return be.rc;
    if (rc < 0) return rc;
    if (rc > 0) return fail(c, SB_ERR_ARG, err);
    if (e != cudaSuccess) return cuda_fail(c, e, "fflonk prove");
    c->last_ms[0] = elapsed(c, 0, 1);
    return 0;
}

}  // namespace
