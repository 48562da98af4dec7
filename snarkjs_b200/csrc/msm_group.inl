// msm_group.inl — the untyped entry points of one (curve, group) pair (msm_entry.h), written once.
// A translation unit defines SB_GROUP (name prefix) and SB_FIELD (coordinate field) and includes this file; the four
// pairs stay in separate .cu files only so that the heavy kernel instantiations compile in parallel.
#include "msm_host.cuh"
#include "msm_entry.h"
#define SB_CAT2(a, b) a##_##b
#define SB_CAT(a, b) SB_CAT2(a, b)
#define SB_FN(name) SB_CAT(SB_GROUP, name)
namespace sb {
typedef SB_FIELD FT;
typedef XYZZ<FT> PT;
int SB_FN(buckets)(const void* d_bases, const MsmSorted& s, MsmScratch& scratch, cudaStream_t stream, void* d_wsum, MsmLaunchStats* stats,
                   cudaStream_t tail_stream, cudaEvent_t ev_acc) {
    return msm_buckets<FT>((const Affine<FT>*)d_bases, s, scratch, stream, (PT*)d_wsum, stats, tail_stream, ev_acc);
}
void SB_FN(combine)(const uint8_t* wsum_host, const MsmGeom& g, uint8_t* acc_xyzz) {
    PT acc; memcpy(&acc, acc_xyzz, sizeof acc);
    const uint32_t NW = g.windows();
    std::vector<PT> ws(NW);
    const WsPlan wp = ws_plan(g);
    if (wp.ok && g_msm_tuning[1] == 0) {
        // k_ws_final's five parts per window: 2^(m+5) a0 + 2^m a1 + 2^5 a2 + a3 + a4
        std::vector<PT> parts((size_t)5 * NW); memcpy(parts.data(), wsum_host, parts.size() * sizeof(PT));
        for (uint32_t w = 0; w < NW; w++) {
            const PT* a = &parts[(size_t)5 * w];
            PT t = a[0]; for (int k = 0; k < 5; k++) t = PT::dbl(t);
            t.add(a[1]); for (int k = 0; k < wp.m; k++) t = PT::dbl(t);
            PT u = a[2]; for (int k = 0; k < 5; k++) u = PT::dbl(u);
            u.add(a[3]); u.add(a[4]); t.add(u);
            ws[w] = t;
        }
    } else memcpy(ws.data(), wsum_host, (size_t)NW * sizeof(PT));   // legacy kernels: one point per window
    if (g.precomp) acc.add(ws[0]);                                   // single shared bucket set: no Horner
    else msm_combine_host<FT>(ws.data(), g, acc);
    memcpy(acc_xyzz, &acc, sizeof acc);
}
void SB_FN(add)(uint8_t* acc_xyzz, const uint8_t* other_xyzz) {
    PT a, b; memcpy(&a, acc_xyzz, sizeof a); memcpy(&b, other_xyzz, sizeof b); a.add(b); memcpy(acc_xyzz, &a, sizeof a);
}
void SB_FN(to_jacobian)(const uint8_t* xyzz, uint8_t* out) { PT p; memcpy(&p, xyzz, sizeof p); xyzz_to_jacobian_bytes<FT>(p, out); }
void SB_FN(to_affine)(const uint8_t* xyzz, uint8_t* out) {
    PT p; memcpy(&p, xyzz, sizeof p);
    if (p.is_inf()) { memset(out, 0, 2 * sizeof(FT)); return; }
    FT x = FT::mul(p.x, FT::inv(p.zz)), y = FT::mul(p.y, FT::inv(p.zzz));
    memcpy(out, &x, sizeof x); memcpy(out + sizeof x, &y, sizeof y);
}
void SB_FN(from_affine)(const uint8_t* aff, uint8_t* xyzz) {
    Affine<FT> a; memcpy(&a, aff, sizeof a);
    PT p = PT::inf();
    if (!a.is_inf()) { p.x = a.x; p.y = a.y; p.zz = FT::one(); p.zzz = FT::one(); }
    memcpy(xyzz, &p, sizeof p);
}
void SB_FN(times)(const uint8_t* xyzz, const uint8_t* k, int nbytes, uint8_t* out) {
    PT p; memcpy(&p, xyzz, sizeof p);
    PT r = PT::inf();
    for (int i = nbytes * 8 - 1; i >= 0; i--) { r = PT::dbl(r); if ((k[i >> 3] >> (i & 7)) & 1) r.add(p); }
    memcpy(out, &r, sizeof r);
}
int SB_FN(gen_points)(const uint8_t* gen_affine, uint64_t seed, uint64_t n, void* d_out, cudaStream_t stream) {
    Affine<FT> g; memcpy(&g, gen_affine, sizeof g);
    if (n) k_gen_points<FT><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(g, seed, n, (Affine<FT>*)d_out);
    return (int)cudaGetLastError();
}
int SB_FN(precompute)(const void* d_bases, uint64_t n, int c, int W, void* d_table, cudaStream_t stream) {
    if (n) k_precompute<FT><<<(unsigned)((n + 127) / 128), 128, 0, stream>>>((const Affine<FT>*)d_bases, n, c, W, (Affine<FT>*)d_table);
    return (int)cudaGetLastError();
}
uint32_t SB_FN(xyzz_bytes)() { return (uint32_t)sizeof(PT); }
}
#undef SB_FN
#undef SB_CAT
#undef SB_CAT2
