// Host-side check of the exact template code the GPU kernels use (fp.cuh / ec.cuh compiled with g++,
// PTX carry chains emulated) against the CPU oracle.  Built and run by tests/test_host_templates.py.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <dlfcn.h>
#include <vector>
#include "../../snarkjs_b200/csrc/ec.cuh"
using namespace sb;

typedef int (*field_op_t)(int, int, const uint8_t*, const uint8_t*, uint8_t*);
typedef int (*field_const_t)(int, int, uint8_t*);
typedef int (*group_op_t)(int, int, int, const uint8_t*, const uint8_t*, uint8_t*);
typedef int (*gen_points_t)(int, int, const uint8_t*, uint64_t, uint64_t, uint8_t*);

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

template <class P> static int check_field(int fid, field_op_t fop, field_const_t fconst, const char* name) {
    typedef Fp<P> F; const int N = P::N;
    uint8_t pb[48]; fconst(fid, 0, pb);
    uint32_t p[12]; memcpy(p, pb, 4 * N);
    for (int i = 0; i < N; i++) if (p[i] != P::p(i)) { printf("%s modulus mismatch limb %d\n", name, i); return 1; }
    int bad = 0;
    for (int it = 0; it < 20000; it++) {
        F a, b;
        for (int i = 0; i < N; i++) { a.v[i] = (uint32_t)rnd(); b.v[i] = (uint32_t)rnd(); }
        // reduce below p by clearing top bits then conditional check (oracle canonicalises via add 0)
        int topbits = 0; { uint32_t t = p[N - 1]; while (t) { topbits++; t >>= 1; } }
        uint32_t mask = (topbits >= 32) ? 0xffffffffu : ((1u << (topbits - 1)) - 1);
        a.v[N - 1] &= mask; b.v[N - 1] &= mask;
        if (it == 0) { a = F::zero(); }
        if (it == 1) { for (int i = 0; i < N; i++) a.v[i] = p[i]; a.v[0] -= 1; b = a; }   // p-1
        if (it == 2) { a = F::zero(); a.v[0] = 1; }
        uint8_t r[48];
        F m = F::mul(a, b); fop(fid, 2, (uint8_t*)a.v, (uint8_t*)b.v, r); if (memcmp(r, m.v, 4 * N)) { bad++; if (bad < 4) printf("%s mul mismatch it=%d\n", name, it); }
        F s = F::add(a, b); fop(fid, 0, (uint8_t*)a.v, (uint8_t*)b.v, r); if (memcmp(r, s.v, 4 * N)) { bad++; if (bad < 4) printf("%s add mismatch it=%d\n", name, it); }
        F d = F::sub(a, b); fop(fid, 1, (uint8_t*)a.v, (uint8_t*)b.v, r); if (memcmp(r, d.v, 4 * N)) { bad++; if (bad < 4) printf("%s sub mismatch it=%d\n", name, it); }
        F n = F::neg(a);    fop(fid, 3, (uint8_t*)a.v, nullptr, r);        if (memcmp(r, n.v, 4 * N)) { bad++; if (bad < 4) printf("%s neg mismatch it=%d\n", name, it); }
        F fm = F::from_mont(a); fop(fid, 6, (uint8_t*)a.v, nullptr, r);    if (memcmp(r, fm.v, 4 * N)) { bad++; if (bad < 4) printf("%s from_mont mismatch it=%d\n", name, it); }
        if constexpr (P::p(N - 1) < 0x55555555u) {   // dual-product multiply vs two oracle multiplies + add
            F c2, d2; for (int i = 0; i < N; i++) { c2.v[i] = (uint32_t)rnd(); d2.v[i] = (uint32_t)rnd(); }
            c2.v[N - 1] &= mask; d2.v[N - 1] &= mask;
            if (it == 3) { for (int i = 0; i < N; i++) { c2.v[i] = p[i]; d2.v[i] = p[i]; } c2.v[0] -= 1; d2.v[0] -= 1; a = c2; b = d2; }
            // canonicalise the random operands through the oracle (add 0)
            uint8_t z0[48] = {0}; uint8_t t1[48], t2[48], t3[48];
            fop(fid, 0, (uint8_t*)c2.v, z0, t1); memcpy(c2.v, t1, 4 * N); fop(fid, 0, (uint8_t*)d2.v, z0, t1); memcpy(d2.v, t1, 4 * N);
            F aa, bb; fop(fid, 0, (uint8_t*)a.v, z0, t1); memcpy(aa.v, t1, 4 * N); fop(fid, 0, (uint8_t*)b.v, z0, t1); memcpy(bb.v, t1, 4 * N);
            F m2 = F::mul2(aa, bb, c2, d2);
            fop(fid, 2, (uint8_t*)aa.v, (uint8_t*)bb.v, t1); fop(fid, 2, (uint8_t*)c2.v, (uint8_t*)d2.v, t2); fop(fid, 0, t1, t2, t3);
            if (memcmp(t3, m2.v, 4 * N)) { bad++; if (bad < 4) printf("%s mul2 mismatch it=%d\n", name, it); }
        }
        if (it < 600) { F ib = F::inv_binary(a); fop(fid, 4, (uint8_t*)a.v, nullptr, r); if (memcmp(r, ib.v, 4 * N)) { bad++; if (bad < 4) printf("%s inv_binary mismatch it=%d\n", name, it); } }
    }
    printf("%s: %s\n", name, bad ? "FAIL" : "ok");
    return bad;
}

// XYZZ -> affine through the oracle: x = X/ZZ, y = Y/ZZZ (inversion via oracle op 4)
template <class P> static void fe_inv(int fid, field_op_t fop, const Fp<P>& a, Fp<P>& out) { fop(fid, 4, (const uint8_t*)a.v, nullptr, (uint8_t*)out.v); }

template <class P> static int check_g1(int curve, int fid, field_op_t fop, field_const_t fconst, group_op_t gop, gen_points_t gen,
                                        const uint8_t* gen_aff, const char* name) {
    typedef Fp<P> F; const int N = P::N; const int n8 = 4 * N;
    F one; fconst(fid, 1, (uint8_t*)one.v);
    const int NP = 64;
    std::vector<uint8_t> pts(NP * 2 * n8);
    gen(curve, 1, gen_aff, 5, NP, pts.data());
    int bad = 0;
    // accumulate all points with XYZZ mixed adds incl. special cases: repeat (doubling), negation (cancel), infinity base
    XYZZ<F> acc = XYZZ<F>::inf();
    std::vector<uint8_t> jac(3 * n8), tmp(3 * n8), aff(2 * n8);
    // oracle accumulator starts at zero
    memset(jac.data(), 0, 3 * n8); memcpy(jac.data() + n8, one.v, n8);
    auto add_pt = [&](const uint8_t* p, bool negate) {
        F x, y; memcpy(x.v, p, n8); memcpy(y.v, p + n8, n8);
        Affine<F> a{x, y};
        if (a.is_inf()) return;
        if (negate) y = F::neg(y);
        acc.add_affine(x, y, one);
        uint8_t pa[2 * 48]; memcpy(pa, x.v, n8); memcpy(pa + n8, y.v, n8);
        gop(curve, 1, 4, jac.data(), pa, tmp.data()); jac = tmp;
    };
    auto compare = [&](const char* what) {
        gop(curve, 1, 2, jac.data(), nullptr, aff.data());
        uint8_t mine[2 * 48];
        if (acc.is_inf()) memset(mine, 0, 2 * n8);
        else { F zi, zzi; fe_inv<P>(fid, fop, acc.zz, zi); fe_inv<P>(fid, fop, acc.zzz, zzi);
               F ax = F::mul(acc.x, zi), ay = F::mul(acc.y, zzi); memcpy(mine, ax.v, n8); memcpy(mine + n8, ay.v, n8); }
        if (memcmp(mine, aff.data(), 2 * n8)) { bad++; printf("%s %s mismatch\n", name, what); }
    };
    for (int i = 0; i < NP; i++) { add_pt(pts.data() + i * 2 * n8, i % 5 == 3); if (i % 7 == 0) compare("running"); }
    compare("sum");
    // doubling path: acc = P ; acc += P
    acc = XYZZ<F>::inf(); memset(jac.data(), 0, 3 * n8); memcpy(jac.data() + n8, one.v, n8);
    add_pt(pts.data(), false); add_pt(pts.data(), false); compare("double");
    add_pt(pts.data() + 2 * n8, false); compare("double+1");
    // cancellation: P + (-P) = inf, then add another
    acc = XYZZ<F>::inf(); memset(jac.data(), 0, 3 * n8); memcpy(jac.data() + n8, one.v, n8);
    add_pt(pts.data(), false); add_pt(pts.data(), true); compare("cancel");
    if (!acc.is_inf()) { bad++; printf("%s cancel not inf\n", name); }
    add_pt(pts.data() + 4 * n8, false); compare("after-cancel");
    // full XYZZ add: (sum of first half) + (sum of second half) == total; plus self-add (doubling) and cancel
    XYZZ<F> h1 = XYZZ<F>::inf(), h2 = XYZZ<F>::inf();
    acc = XYZZ<F>::inf(); memset(jac.data(), 0, 3 * n8); memcpy(jac.data() + n8, one.v, n8);
    for (int i = 0; i < NP; i++) {
        F x, y; memcpy(x.v, pts.data() + i * 2 * n8, n8); memcpy(y.v, pts.data() + i * 2 * n8 + n8, n8);
        (i < NP / 2 ? h1 : h2).add_affine(x, y, one);
        add_pt(pts.data() + i * 2 * n8, false);
    }
    XYZZ<F> tot = h1; tot.add(h2); acc = tot; compare("xyzz-add");
    XYZZ<F> d1 = h1; d1.add(h1); XYZZ<F> d2 = XYZZ<F>::dbl(h1);
    { F l = F::mul(d1.x, d2.zz), r = F::mul(d2.x, d1.zz); if (!(l == r)) { bad++; printf("%s xyzz self-add != dbl\n", name); } }
    XYZZ<F> ng = h1; ng.y = F::neg(ng.y); XYZZ<F> z = h1; z.add(ng); if (!z.is_inf()) { bad++; printf("%s xyzz cancel\n", name); }
    printf("%s: %s\n", name, bad ? "FAIL" : "ok");
    return bad;
}

// Fp2 multiply (dual-product schoolbook) against Karatsuba built from single multiplies, and squaring against mul
template <class P> static int check_fp2(const char* name) {
    typedef Fp<P> B; typedef Fp2<P> F2; const int N = P::N;
    int bad = 0;
    for (int it = 0; it < 5000; it++) {
        F2 x, y; B* parts[4] = {&x.a, &x.b, &y.a, &y.b};
        for (B* q : parts) { for (int i = 0; i < N; i++) q->v[i] = (uint32_t)rnd(); q->v[N - 1] &= 0x0fffffffu; *q = B::add(*q, B::zero()); }
        if (it == 0) { x.a = B::zero(); }
        if (it == 1) { for (B* q : parts) { for (int i = 0; i < N; i++) q->v[i] = P::p(i); q->v[0] -= 1; } }
        F2 m = F2::mul_i(x, y);
        B A = B::mul(x.a, y.a), Bb = B::mul(x.b, y.b), C = B::mul(B::add(x.a, x.b), B::add(y.a, y.b));
        B c0 = B::sub(A, Bb), c1 = B::sub(B::sub(C, A), Bb);
        if (!(m.a == c0) || !(m.b == c1)) { bad++; if (bad < 4) printf("%s fp2 mul mismatch it=%d\n", name, it); }
        F2 ml = F2::mul_lazy(x, y);
        if (!(ml == m)) { bad++; if (bad < 4) printf("%s fp2 mul_lazy mismatch it=%d\n", name, it); }
        { uint32_t T[2 * N]; B::mul_wide(x.a.v, y.b.v, T); B rr = B::redc_wide(T); if (!(rr == B::mul(x.a, y.b))) { bad++; if (bad < 4) printf("%s mul_wide/redc mismatch it=%d\n", name, it); } }
        F2 s = F2::sqr_i(x), s2 = F2::mul_i(x, x);
        if (!(s == s2)) { bad++; if (bad < 4) printf("%s fp2 sqr mismatch it=%d\n", name, it); }
    }
    printf("%s: %s\n", name, bad ? "FAIL" : "ok");
    return bad;
}

int main(int argc, char** argv) {
    void* h = dlopen(argc > 1 ? argv[1] : "oracle/liboracle.so", RTLD_NOW);
    if (!h) { printf("dlopen failed: %s\n", dlerror()); return 2; }
    field_op_t fop = (field_op_t)dlsym(h, "or_field_op");
    field_const_t fconst = (field_const_t)dlsym(h, "or_field_const");
    group_op_t gop = (group_op_t)dlsym(h, "or_group_op");
    gen_points_t gen = (gen_points_t)dlsym(h, "or_gen_points");
    ((int (*)())dlsym(h, "or_init"))();
    int bad = 0;
    bad += check_field<BnFq>(0, fop, fconst, "BnFq");
    bad += check_field<BnFr>(1, fop, fconst, "BnFr");
    bad += check_field<BlsFq>(2, fop, fconst, "BlsFq");
    bad += check_field<BlsFr>(3, fop, fconst, "BlsFr");
    // generators in Montgomery affine form are passed on argv as hex? simpler: derive from oracle toMont of (1,2)
    { uint8_t g[64] = {0}, gm[64]; g[0] = 1; g[32] = 2; fop(0, 5, g, nullptr, gm); fop(0, 5, g + 32, nullptr, gm + 32);
      bad += check_g1<BnFq>(0, 0, fop, fconst, gop, gen, gm, "BN254 G1 XYZZ"); }
    bad += check_fp2<BnFq>("BN254 Fq2");
    bad += check_fp2<BlsFq>("BLS12-381 Fq2");
    printf(bad ? "HOST CHECK FAILED\n" : "HOST CHECK PASSED\n");
    return bad ? 1 : 0;
}
