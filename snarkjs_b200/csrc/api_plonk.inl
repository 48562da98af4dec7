// api_plonk.inl — PLONK proving on the device (part of api.cu's translation unit: it uses sb_ctx and its helpers).
//
//   sb_plonk_load     zkey sections 3-14 -> HBM (selector / sigma / Lagrange coefficient and 4n-evaluation arrays, wire maps,
//                     additions sorted by dependency level, the PTau bases with their window tables), work buffers
//   sb_plonk_prove    plonk_flow.h's five rounds (src/plonk_prove.js:47-889) on the CUDA backend below: the kernels of
//                     plonk.cuh, cub scans/reductions over field elements, the NTT passes (ntt.cuh) and the MSM pipeline
//                     (msm.cuh, table mode).  Only the transcript hashing and ~40 scalar operations run on the host.
//
// No CPU fallback: every bulk step is a kernel launch on c->stream.
#include <chrono>
#include <cub/cub.cuh>
#include "plonk_flow.h"

namespace {

struct PlonkKeyDev {
    PlonkZkey z;                                  // header values; pointers into hdr
    std::vector<uint8_t> hdr;                     // copy of the header section
    std::vector<void*> allocs;
    uint32_t* d_add_sig = nullptr; void* d_add_fac = nullptr; uint32_t* d_add_order = nullptr; std::vector<uint32_t> level_end;
    uint32_t* d_map[3] = {nullptr, nullptr, nullptr};
    void* d_q_coef[5] = {nullptr}; void* d_q_ev[5] = {nullptr}; void* d_s_coef[3] = {nullptr}; void* d_s_ev[3] = {nullptr}; void* d_lag = nullptr;
    void* d_ptau = nullptr; void* t_ptau = nullptr; MsmGeom gp{};
    void *d_wlo = nullptr, *d_whi = nullptr, *d_w4lo = nullptr, *d_w4hi = nullptr; int wh = 0, w4h = 0;
    void* d_pow[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}}; uint64_t pow_nhi = 0; int pow_h = 0;
    void* work[27] = {nullptr};              // W | 8 x n | 10 x (n + 8) | 8 x 4n   (order of PlonkWork)
    void* d_cub = nullptr; size_t cub_bytes = 0;
    int* d_flag = nullptr; void* d_red = nullptr;
    std::vector<uint8_t> host_pub; uint64_t n_wit_resident = 0;   // witness[0..nPublic] on the host + length of the witness left in work[0] by the last proof
    void* commit_scratch() const { return work[18]; }   // PlonkWork::scal: free whenever a Montgomery polynomial is committed
};

void plonk_free_key(PlonkKeyDev* k) {
    for (void* p : k->allocs) if (p) cudaFree(p);
    if (k->t_ptau) cudaFree(k->t_ptau);
    delete k;
}

// K = the device-side key (PlonkKeyDev here, FflonkKeyDev in api_fflonk.inl): both expose d_flag, d_red, d_cub, cub_bytes,
// d_ptau, t_ptau, gp, d_pow, pow_h, pow_nhi and commit_scratch()
template <class F, class K = PlonkKeyDev> struct CudaPlonkBackend {
    sb_ctx* c; K* key; int rc = 0;
    const void* resident_dst = nullptr;   // set: the witness already sits in this buffer (sb_*_prove_resident), its upload is skipped
    std::chrono::steady_clock::time_point t_mark = std::chrono::steady_clock::now();
    // end of round r: host wall clock since the previous mark (each round ends on a synchronising commit) -> sb_last_ms(r)
    void mark(int r) {
        auto now = std::chrono::steady_clock::now();
        if (r >= 1 && r <= 5) c->last_ms[r] = std::chrono::duration<float, std::milli>(now - t_mark).count();
        t_mark = now;
    }
    cudaStream_t st() const { return c->stream; }
    void note(cudaError_t e, const char* what) { if (e != cudaSuccess && !rc) rc = cuda_fail(c, e, what); }
    void launched(const char* what) { c->launches++; note(cudaGetLastError(), what); }
    static unsigned grid(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

    void upload(F* dst, const F* host, size_t n) { if (dst == resident_dst) return; note(h2d(c, dst, host, n * sizeof(F)), "plonk upload"); }
    void download(F* host, const F* src, size_t n) { note(d2h(c, host, src, n * sizeof(F)), "plonk download"); }
    void zero(F* p, size_t n) { if (n) note(cudaMemsetAsync(p, 0, n * sizeof(F), st()), "plonk memset"); }
    void copy(F* dst, const F* src, size_t n) { if (n) note(cudaMemcpyAsync(dst, src, n * sizeof(F), cudaMemcpyDeviceToDevice, st()), "plonk copy"); }
    F* ntt(F* a, F* b, uint64_t n, bool inverse) {
        void* res = a;
        int r = ntt_dev(c, a, b, n, inverse ? 1 : 0, nullptr, true, &res);
        if (r && !rc) rc = r;
        return (F*)res;
    }
    int read_flag() {
        int f = 0;
        note(cudaMemcpyAsync(&f, key->d_flag, sizeof f, cudaMemcpyDeviceToHost, st()), "plonk flag");
        note(cudaStreamSynchronize(st()), "plonk flag");
        note(cudaMemsetAsync(key->d_flag, 0, sizeof(int), st()), "plonk flag");
        return rc ? 0 : f;
    }
    int commit_plain(const F* scal, uint64_t len, uint8_t* affine) {
        if (rc) return rc;
        std::vector<uint8_t> acc(c->g1.xyzz_bytes, 0);
        int r = key->t_ptau ? msm_dev_accumulate(c, c->g1, key->t_ptau, (const uint8_t*)scal, 32, len, acc.data(), &key->gp, 0)
                            : msm_dev_accumulate(c, c->g1, key->d_ptau, (const uint8_t*)scal, 32, len, acc.data());
        if (r) return r;
        c->g1.to_affine(acc.data(), affine);
        return 0;
    }
    int commit(const F* coef, uint64_t len, uint8_t* affine) {
        if (rc) return rc;
        F* scal = (F*)key->commit_scratch();
        int r = fr_convert(c->curve, coef, scal, len, 0, st()); c->launches++;
        if (r) return cuda_fail(c, (cudaError_t)r, "fr_convert");
        return commit_plain(scal, len, affine);
    }
    void additions(const PlonkKeyView<F>& k, F* W) {
        uint32_t lo = 0;
        for (uint32_t hi : k.level_end) {
            if (hi > lo) { k_pl_additions<F><<<grid(hi - lo, 128), 128, 0, st()>>>(k.add_order, lo, hi, k.add_sig, k.add_fac, W, k.nVars - k.nAdditions, k.nVars); launched("k_pl_additions"); }
            lo = hi;
        }
    }
    void wires(const PlonkKeyView<F>& k, const F* W, F* A, F* B, F* C) {
        PlonkMaps mp; mp.m[0] = k.map[0]; mp.m[1] = k.map[1]; mp.m[2] = k.map[2]; mp.out[0] = A; mp.out[1] = B; mp.out[2] = C;
        dim3 g(grid(k.n, 256), 3);
        k_pl_wires<F><<<g, 256, 0, st()>>>(mp, W, k.nVars, k.nConstraints, k.n); launched("k_pl_wires");
    }
    void blind(F* p, uint64_t n, const F* bf, int cnt) {
        PlonkBlind<F> b; b.cnt = cnt; for (int i = 0; i < 3; i++) b.bf[i] = i < cnt ? bf[i] : F::zero();
        k_pl_blind<F><<<1, 32, 0, st()>>>(p, n, b); launched("k_pl_blind");
    }
    int z(const PlonkKeyView<F>& k, const PlonkRound<F>& r, PlonkWork<F>& w) {
        const uint64_t n = k.n;
        k_pl_z_terms<F><<<grid(n, 128), 128, 0, st()>>>(n, w.bufA, w.bufB, w.bufC, k.s_ev[0], k.s_ev[1], k.s_ev[2], k.wpow, r, w.num, w.den); launched("k_pl_z_terms");
        k_pl_ratio<F><<<grid((n + PL_INV_CHUNK - 1) / PL_INV_CHUNK, 64), 64, 0, st()>>>(w.den, w.num, w.ratio, n); launched("k_pl_ratio");
        size_t bytes = key->cub_bytes;
        note(cub::DeviceScan::ExclusiveScan(key->d_cub, bytes, (const F*)w.ratio, w.bufZ, FrMulOp(), F::one(), (int)n, st()), "cub ExclusiveScan"); c->launches += 2;
        k_pl_z_check<F><<<1, 32, 0, st()>>>(w.bufZ, w.ratio, n, key->d_flag); launched("k_pl_z_check");
        return read_flag();
    }
    void t(const PlonkKeyView<F>& k, const PlonkRound<F>& r, PlonkWork<F>& w) {
        PlonkTIn in;
        in.A = w.evA; in.B = w.evB; in.C = w.evC; in.Z = w.evZ;
        in.QM = k.q_ev[0]; in.QL = k.q_ev[1]; in.QR = k.q_ev[2]; in.QO = k.q_ev[3]; in.QC = k.q_ev[4];
        in.S1 = k.s_ev[0]; in.S2 = k.s_ev[1]; in.S3 = k.s_ev[2]; in.LAG = k.lag; in.pubA = w.bufA; in.n_public = k.nPublic;
        const uint64_t n4 = 4ull * k.n;
        k_pl_t<F><<<grid(n4, 128), 128, 0, st()>>>(n4, in, k.w4pow, r, w.T, w.Tz); launched("k_pl_t");
    }
    int divzh(uint64_t n, const F* t, const F* tz, F* out) {
        k_pl_divzh<F><<<grid(n, 128), 128, 0, st()>>>(n, t, tz, out, key->d_flag); launched("k_pl_divzh");
        return read_flag();
    }
    void tsplit(uint64_t n, const F* t, const F& b10, const F& b11, F* T1, F* T2, F* T3) {
        PlonkB2<F> b; b.b10 = b10; b.b11 = b11;
        k_pl_tsplit<F><<<grid(n + 6, 256), 256, 0, st()>>>(n, t, b, T1, T2, T3); launched("k_pl_tsplit");
    }
    void make_pow(const F& base, uint64_t count, PlonkPow<F>& out, int slot) {
        std::vector<F> lo, hi;
        plonk_pow_tables<F>(base, key->pow_h, key->pow_nhi, lo, hi);
        (void)count;
        note(cudaMemcpyAsync(key->d_pow[slot][0], lo.data(), lo.size() * sizeof(F), cudaMemcpyHostToDevice, st()), "plonk pow upload");
        note(cudaMemcpyAsync(key->d_pow[slot][1], hi.data(), hi.size() * sizeof(F), cudaMemcpyHostToDevice, st()), "plonk pow upload");
        note(cudaStreamSynchronize(st()), "plonk pow upload");      // lo / hi are stack vectors
        out.lo = (const F*)key->d_pow[slot][0]; out.hi = (const F*)key->d_pow[slot][1]; out.h = key->pow_h;
    }
    F eval(const F* f, uint64_t len, const PlonkPow<F>& pw, F* g, F*) {
        PlonkOne<F> z0; z0.x = F::zero();
        k_pl_mul_pow<F><<<grid(len, 256), 256, 0, st()>>>(f, len, len, pw, z0, g); launched("k_pl_mul_pow");
        size_t bytes = key->cub_bytes;
        note(cub::DeviceReduce::Reduce(key->d_cub, bytes, (const F*)g, (F*)key->d_red, (int)len, FrAddOp(), F::zero(), st()), "cub Reduce"); c->launches += 2;
        F out = F::zero();
        note(cudaMemcpyAsync(&out, key->d_red, sizeof(F), cudaMemcpyDeviceToHost, st()), "plonk eval");
        note(cudaStreamSynchronize(st()), "plonk eval");
        return out;
    }
    int quotient(const F* f, const PlonkLinIn* lin, const PlonkLin<F>* L, uint64_t n, uint64_t len, uint64_t m, const F& sub0,
                 const PlonkPow<F>& pw, const PlonkPow<F>& ipw, F* g, F* P, F* q_plain) {
        if (lin) { k_pl_wxi<F><<<grid(n + 6, 128), 128, 0, st()>>>(n, *lin, *L, pw, g); launched("k_pl_wxi"); }
        else { PlonkOne<F> s0; s0.x = sub0; k_pl_mul_pow<F><<<grid(m, 256), 256, 0, st()>>>(f, len, m, pw, s0, g); launched("k_pl_mul_pow"); }
        size_t bytes = key->cub_bytes;
        note(cub::DeviceScan::InclusiveScan(key->d_cub, bytes, (const F*)g, P, FrAddOp(), (int)m, st()), "cub InclusiveScan"); c->launches += 2;
        k_pl_quot<F><<<grid(m, 256), 256, 0, st()>>>(m, P, ipw, q_plain, key->d_flag); launched("k_pl_quot");
        return read_flag();
    }
};

template <class F> cudaError_t plonk_cub_bytes(uint64_t n, size_t* out) {
    size_t a = 0, b = 0, d = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveScan(nullptr, a, (const F*)nullptr, (F*)nullptr, FrMulOp(), F::one(), (int)(n + PLONK_PAD), (cudaStream_t)0);
    if (e != cudaSuccess) return e;
    e = cub::DeviceScan::InclusiveScan(nullptr, b, (const F*)nullptr, (F*)nullptr, FrAddOp(), (int)(n + PLONK_PAD), (cudaStream_t)0);
    if (e != cudaSuccess) return e;
    e = cub::DeviceReduce::Reduce(nullptr, d, (const F*)nullptr, (F*)nullptr, (int)(n + PLONK_PAD), FrAddOp(), F::zero(), (cudaStream_t)0);
    if (e != cudaSuccess) return e;
    *out = std::max(a, std::max(b, d)) + 256;
    return cudaSuccess;
}

template <class PR> int plonk_load_impl(sb_ctx* c, const uint8_t* zkey, uint64_t zlen, uint64_t* handle) {
    typedef Fp<PR> F;
    PlonkKeyDev* k = new PlonkKeyDev();
    std::string err;
    PlonkZkey z;
    if (plonk_parse_zkey(zkey, zlen, z, err)) { delete k; return fail(c, SB_ERR_FORMAT, err); }
    if (!modulus_matches(z.q, z.n8q, c->curve, false) || !modulus_matches(z.r, z.n8r, c->curve, true)) { delete k; return fail(c, SB_ERR_ARG, "zkey curve does not match the context curve"); }
    if (z.power + 2 > c->fr_s) { delete k; return fail(c, SB_ERR_ARG, "domain too large for the 2-adicity of Fr"); }
    // keep the header (k1, k2, the eight commitments) on the host
    k->hdr.assign(z.sec[2].p, z.sec[2].p + z.sec[2].len);
    k->z = z;
    { const ptrdiff_t d = k->hdr.data() - z.sec[2].p; k->z.q += d; k->z.r += d; k->z.k1 += d; k->z.k2 += d; k->z.hdr_pts += d; k->z.X_2 += d; }
    for (auto& s : k->z.sec) s = PlonkZkey::Sec();
    const uint64_t n = z.n, sd = n * 32;
    bool ok = true;
    auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (!ok) return nullptr; if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) { cudaGetLastError(); ok = false; return nullptr; } k->allocs.push_back(p); return p; };
    auto put = [&](void* dst, const void* src, size_t bytes) { if (ok && bytes && h2d(c, dst, src, bytes) != cudaSuccess) ok = false; };
    // additions: (s1, s2) and (f1, f2) arrays + level order
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
    for (int j = 0; j < 3; j++) { k->d_s_coef[j] = dalloc(sd); k->d_s_ev[j] = dalloc(4 * sd); put(k->d_s_coef[j], z.sec[12].p + 5 * sd * j, sd); put(k->d_s_ev[j], z.sec[12].p + 5 * sd * j + sd, 4 * sd); }
    {
        const uint32_t nl = z.nPublic > 1 ? z.nPublic : 1, have = (uint32_t)(z.sec[13].len / (5 * sd));
        k->d_lag = dalloc((size_t)nl * 4 * sd);
        if (ok) cudaMemsetAsync(k->d_lag, 0, (size_t)nl * 4 * sd, c->stream);
        for (uint32_t j = 0; j < nl && j < have; j++) put((uint8_t*)k->d_lag + (size_t)j * 4 * sd, z.sec[13].p + 5 * sd * j + sd, 4 * sd);
    }
    const uint64_t npts = n + 6;
    k->d_ptau = dalloc(npts * c->g1.aff_bytes); put(k->d_ptau, z.sec[14].p, npts * c->g1.aff_bytes);
    // powers of w_n and w_4n (two-level tables), per-proof power tables for xi, xi w, and their inverses
    {
        F wn, w4; memcpy(&wn, c->roots[z.power].data(), 32); memcpy(&w4, c->roots[z.power + 2].data(), 32);
        std::vector<F> lo, hi;
        k->wh = plonk_pow_h(n); plonk_pow_tables<F>(wn, k->wh, (n >> k->wh) + 1, lo, hi);
        k->d_wlo = dalloc(lo.size() * 32); k->d_whi = dalloc(hi.size() * 32); put(k->d_wlo, lo.data(), lo.size() * 32); put(k->d_whi, hi.data(), hi.size() * 32);
        if (ok) cudaStreamSynchronize(c->stream);
        k->w4h = plonk_pow_h(4 * n); plonk_pow_tables<F>(w4, k->w4h, ((4 * n) >> k->w4h) + 1, lo, hi);
        k->d_w4lo = dalloc(lo.size() * 32); k->d_w4hi = dalloc(hi.size() * 32); put(k->d_w4lo, lo.data(), lo.size() * 32); put(k->d_w4hi, hi.data(), hi.size() * 32);
        if (ok) cudaStreamSynchronize(c->stream);
        k->pow_h = plonk_pow_h(n + PLONK_PAD); k->pow_nhi = ((n + PLONK_PAD) >> k->pow_h) + 1;
        for (int s = 0; s < 4; s++) { k->d_pow[s][0] = dalloc(((size_t)1 << k->pow_h) * 32); k->d_pow[s][1] = dalloc(k->pow_nhi * 32); }
    }
    // work buffers: W | 8 x n | 10 x (n + 8) | 8 x 4n
    k->work[0] = dalloc(((size_t)z.nVars + 2) * 32);
    for (int i = 1; i <= 8; i++) k->work[i] = dalloc(sd);
    for (int i = 9; i <= 18; i++) k->work[i] = dalloc(sd + PLONK_PAD * 32);
    for (int i = 19; i <= 26; i++) k->work[i] = dalloc(4 * sd);
    k->d_flag = (int*)dalloc(16); k->d_red = dalloc(64);
    if (ok) { cudaMemsetAsync(k->d_flag, 0, 16, c->stream);
        if (plonk_cub_bytes<F>(n, &k->cub_bytes) != cudaSuccess) ok = false; else k->d_cub = dalloc(k->cub_bytes); }
    if (ok && cudaStreamSynchronize(c->stream) != cudaSuccess) ok = false;
    if (!ok) { cudaGetLastError(); plonk_free_key(k); return fail(c, SB_ERR_NOMEM, "plonk key does not fit in device memory"); }
    if (want_precomp(c, npts)) {
        int rc = build_table(c, c->g1, k->d_ptau, npts, &k->t_ptau, &k->gp);
        if (rc) { plonk_free_key(k); return rc; }
    }
    c->plonk_keys.push_back(k);
    *handle = c->plonk_keys.size();
    return 0;
}

template <class PQ, class PR> int plonk_prove_impl(sb_ctx* c, PlonkKeyDev* kd, const uint8_t* witness, uint64_t n_witness, const uint8_t* blinders, uint8_t* proof) {
    typedef Fp<PR> F;
    const PlonkZkey& z = kd->z;
    PlonkKeyView<F> k;
    k.nVars = z.nVars; k.nPublic = z.nPublic; k.n = z.n; k.nAdditions = z.nAdditions; k.nConstraints = z.nConstraints; k.power = z.power;
    memcpy(&k.k1, z.k1, 32); memcpy(&k.k2, z.k2, 32);
    memcpy(&k.wn, c->roots[z.power].data(), 32); memcpy(&k.w4n, c->roots[z.power + 2].data(), 32);
    { F w2; memcpy(&w2, c->roots[2].data(), 32); plonk_mulz_tables<F>(w2, k.z1, k.z2, k.z3); }
    k.hdr_pts = z.hdr_pts; k.aff_bytes = c->g1.aff_bytes;
    k.add_sig = kd->d_add_sig; k.add_fac = (const F*)kd->d_add_fac; k.add_order = kd->d_add_order; k.level_end = kd->level_end;
    for (int j = 0; j < 3; j++) { k.map[j] = kd->d_map[j]; k.s_coef[j] = (const F*)kd->d_s_coef[j]; k.s_ev[j] = (const F*)kd->d_s_ev[j]; }
    for (int j = 0; j < 5; j++) { k.q_coef[j] = (const F*)kd->d_q_coef[j]; k.q_ev[j] = (const F*)kd->d_q_ev[j]; }
    k.lag = (const F*)kd->d_lag;
    k.wpow.lo = (const F*)kd->d_wlo; k.wpow.hi = (const F*)kd->d_whi; k.wpow.h = kd->wh;
    k.w4pow.lo = (const F*)kd->d_w4lo; k.w4pow.hi = (const F*)kd->d_w4hi; k.w4pow.h = kd->w4h;
    PlonkWork<F> w;
    {
        int wi = 0; auto nx = [&]() { return (F*)kd->work[wi++]; };
        w.W = nx();
        w.bufA = nx(); w.bufB = nx(); w.bufC = nx(); w.bufZ = nx(); w.num = nx(); w.den = nx(); w.ratio = nx(); w.sn = nx();
        w.cA = nx(); w.cB = nx(); w.cC = nx(); w.cZ = nx(); w.T1 = nx(); w.T2 = nx(); w.T3 = nx(); w.g = nx(); w.P = nx(); w.scal = nx();
        w.evA = nx(); w.evB = nx(); w.evC = nx(); w.evZ = nx(); w.T = nx(); w.Tz = nx(); w.s4a = nx(); w.s4b = nx();
    }
    CudaPlonkBackend<F> be; be.c = c; be.key = kd;
    std::string err;
    const size_t pub_bytes = ((size_t)z.nPublic + 1) * 32;
    if (!witness) { witness = kd->host_pub.data(); n_witness = kd->n_wit_resident; be.resident_dst = w.W; }   // resident: only the public signals are read on the host
    else { kd->n_wit_resident = 0; if (n_witness > z.nPublic) kd->host_pub.assign(witness, witness + pub_bytes); }
    tick(c, 0);
    prof_begin(c);
    int rc = plonk_prove_flow<PQ, PR>(be, k, w, witness, n_witness, blinders, proof, err);
    tick(c, 1);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    prof_end(c);
    if (!be.rc && rc == 0 && e == cudaSuccess) kd->n_wit_resident = n_witness;
    if (be.rc) return be.rc;                                   // a CUDA failure underneath explains whatever the flow reported
    if (rc < 0) return rc;                                     // backend error, message already set
    if (rc > 0) return fail(c, SB_ERR_ARG, err);               // the reference's own Error text
    if (e != cudaSuccess) return cuda_fail(c, e, "plonk prove");
    c->last_ms[0] = elapsed(c, 0, 1);
    return 0;
}

}  // namespace
