// Host-side check of the lane-pair extension-field arithmetic (Fp2L, ec.cuh) that k_accumulate_pair runs on the GPU:
// the two lanes of a pair are two host threads executing the very same template code in lockstep; the three pair
// primitives (parity, component exchange, vote) go through a two-slot mailbox with a barrier.  Every result is compared
// with the plain Fp2 / XYZZ<Fp2> code (which tests/host/host_fp_check.cpp pins to the oracle).
// Built and run by tests/test_host_templates.py:  g++ -O2 -std=c++17 -pthread -DSB_PAIR_HOST_EMULATE
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include "../../snarkjs_b200/csrc/ec.cuh"
using namespace sb;

// ---- the pair: a sense-reversing spin barrier and a mailbox --------------------------------------------------------
static std::atomic<int> g_count{0}; static std::atomic<int> g_sense{0};
static thread_local int t_lane = 0; static thread_local int t_sense = 0;
static uint32_t g_box[2][16]; static int g_vote[2];
static void pair_barrier() {
    t_sense ^= 1;
    if (g_count.fetch_add(1) == 1) { g_count.store(0); g_sense.store(t_sense); }
    else while (g_sense.load() != t_sense) std::this_thread::yield();
}
namespace sb {
bool sb_pair_odd() { return t_lane != 0; }
void sb_pair_exchange(const uint32_t* mine, uint32_t* others, int n) {
    memcpy(g_box[t_lane], mine, 4 * n); pair_barrier();
    memcpy(others, g_box[t_lane ^ 1], 4 * n); pair_barrier();
}
bool sb_pair_all(bool p) { g_vote[t_lane] = p; pair_barrier(); bool r = g_vote[0] && g_vote[1]; pair_barrier(); return r; }
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
template <class P> static Fp<P> rand_fe() { Fp<P> a; for (int i = 0; i < P::N; i++) a.v[i] = (uint32_t)rnd(); a.v[P::N - 1] &= 0x0fffffffu; return Fp<P>::add(a, Fp<P>::zero()); }

template <class P> static Fp2L<P> lane_of(const Fp2<P>& x, int lane) { Fp2L<P> r; r.m = lane ? x.b : x.a; return r; }
template <class P> static XYZZ<Fp2L<P>> lane_of(const XYZZ<Fp2<P>>& p, int lane) {
    XYZZ<Fp2L<P>> r; r.x = lane_of(p.x, lane); r.y = lane_of(p.y, lane); r.zz = lane_of(p.zz, lane); r.zzz = lane_of(p.zzz, lane); return r;
}
template <class P> static void join(XYZZ<Fp2<P>>& out, const XYZZ<Fp2L<P>>& l, int lane) {
    (lane ? out.x.b : out.x.a) = l.x.m; (lane ? out.y.b : out.y.a) = l.y.m; (lane ? out.zz.b : out.zz.a) = l.zz.m; (lane ? out.zzz.b : out.zzz.a) = l.zzz.m;
}
template <class T> static bool same(const T& a, const T& b) { return memcmp(&a, &b, sizeof(T)) == 0; }

// a point on y^2 = x^3 + b is not needed: the XYZZ formulas are polynomial identities, so arbitrary (x, y) pairs exercise
// the same code paths; the special cases are forced explicitly (same x and y -> doubling, same x and -y -> cancellation).
template <class P> static int check(const char* name) {
    typedef Fp2<P> F2; typedef Fp2L<P> FL;
    const int NOPS = 400;
    struct Op { int kind; F2 x, y; };
    std::vector<Op> ops(NOPS);
    std::vector<F2> mx(NOPS), my(NOPS);
    for (int i = 0; i < NOPS; i++) {
        ops[i].kind = 0; ops[i].x.a = rand_fe<P>(); ops[i].x.b = rand_fe<P>(); ops[i].y.a = rand_fe<P>(); ops[i].y.b = rand_fe<P>();
        mx[i].a = rand_fe<P>(); mx[i].b = rand_fe<P>(); my[i].a = rand_fe<P>(); my[i].b = rand_fe<P>();
        if (i % 37 == 5) { mx[i].b = Fp<P>::zero(); }
        if (i % 41 == 7) { mx[i].a = Fp<P>::zero(); my[i].b = Fp<P>::zero(); }
    }
    ops[60].kind = 3;    // repeat the previous base right after a reset (acc = P, then + P: doubling)
    ops[70].kind = 4;    // previous base negated after a reset (acc = P, then - P: infinity)
    ops[80].kind = 5;    // base at infinity (all-zero coordinates): skipped by the caller's guard, checked through is_zero
    ops[90].kind = 6;    // y = 0 with the same x after a reset: dbl_affine returns infinity
    // reference run on plain Fp2
    std::vector<XYZZ<F2>> ref(NOPS); std::vector<F2> rmul(NOPS), rsqr(NOPS); std::vector<int> rzero(NOPS);
    {
        const F2 one = F2::one(); XYZZ<F2> acc = XYZZ<F2>::inf();
        for (int i = 0; i < NOPS; i++) {
            F2 px = ops[i].x, py = ops[i].y;
            if (ops[i].kind == 3) { acc = XYZZ<F2>::inf(); acc.add_affine(ops[i - 1].x, ops[i - 1].y, one); px = ops[i - 1].x; py = ops[i - 1].y; }
            if (ops[i].kind == 4) { acc = XYZZ<F2>::inf(); acc.add_affine(ops[i - 1].x, ops[i - 1].y, one); px = ops[i - 1].x; py = F2::neg(ops[i - 1].y); }
            if (ops[i].kind == 5) { px = F2::zero(); py = F2::zero(); }
            if (ops[i].kind == 6) { acc = XYZZ<F2>::inf(); px = ops[i - 1].x; py = F2::zero(); acc.add_affine(px, py, one); }
            if (!(px.is_zero() & py.is_zero())) acc.add_affine(px, F2::cneg(py, (i & 3) == 1), one);
            ref[i] = acc;
            rmul[i] = F2::mul_i(mx[i], my[i]); rsqr[i] = F2::sqr_i(mx[i]); rzero[i] = mx[i].is_zero() ? 1 : 0;
        }
    }
    // lane-pair run: two threads in lockstep
    std::vector<XYZZ<F2>> got(NOPS); std::vector<F2> gmul(NOPS), gsqr(NOPS); std::vector<int> gzero[2]; gzero[0].resize(NOPS); gzero[1].resize(NOPS);
    auto lane_fn = [&](int lane) {
        t_lane = lane; t_sense = 0;
        const FL one = FL::one(); XYZZ<FL> acc = XYZZ<FL>::inf();
        for (int i = 0; i < NOPS; i++) {
            FL px = lane_of(ops[i].x, lane), py = lane_of(ops[i].y, lane);
            if (ops[i].kind == 3) { acc = XYZZ<FL>::inf(); acc.add_affine(lane_of(ops[i - 1].x, lane), lane_of(ops[i - 1].y, lane), one); px = lane_of(ops[i - 1].x, lane); py = lane_of(ops[i - 1].y, lane); }
            if (ops[i].kind == 4) { acc = XYZZ<FL>::inf(); acc.add_affine(lane_of(ops[i - 1].x, lane), lane_of(ops[i - 1].y, lane), one); px = lane_of(ops[i - 1].x, lane); py = FL::neg(lane_of(ops[i - 1].y, lane)); }
            if (ops[i].kind == 5) { px = FL::zero(); py = FL::zero(); }
            if (ops[i].kind == 6) { acc = XYZZ<FL>::inf(); px = lane_of(ops[i - 1].x, lane); py = FL::zero(); acc.add_affine(px, py, one); }
            if (!(px.is_zero() & py.is_zero())) acc.add_affine(px, FL::cneg(py, (i & 3) == 1), one);
            join(got[i], acc, lane);
            FL m = FL::mul_i(lane_of(mx[i], lane), lane_of(my[i], lane)), s = FL::sqr_i(lane_of(mx[i], lane));
            (lane ? gmul[i].b : gmul[i].a) = m.m; (lane ? gsqr[i].b : gsqr[i].a) = s.m;
            gzero[lane][i] = lane_of(mx[i], lane).is_zero() ? 1 : 0;
        }
    };
    g_count = 0; g_sense = 0;
    std::thread t1(lane_fn, 1); lane_fn(0); t1.join();
    int bad = 0, ninf = 0;
    for (int i = 0; i < NOPS; i++) {
        if (!same(ref[i], got[i])) { bad++; if (bad < 5) printf("%s: accumulator mismatch after op %d (kind %d)\n", name, i, ops[i].kind); }
        if (!same(rmul[i], gmul[i])) { bad++; if (bad < 5) printf("%s: mul mismatch %d\n", name, i); }
        if (!same(rsqr[i], gsqr[i])) { bad++; if (bad < 5) printf("%s: sqr mismatch %d\n", name, i); }
        if (rzero[i] != gzero[0][i] || rzero[i] != gzero[1][i]) { bad++; if (bad < 5) printf("%s: is_zero mismatch %d\n", name, i); }
        if (ref[i].is_inf()) ninf++;
    }
    if (!ref[70].is_inf() || !ref[90].is_inf()) { bad++; printf("%s: the forced cancellation / y = 0 cases did not produce infinity\n", name); }
    printf("%s: %s (%d ops, %d at infinity)\n", name, bad ? "FAIL" : "ok", NOPS, ninf);
    return bad;
}

int main() {
    int bad = 0;
    bad += check<BnFq>("BN254 Fq2 lane pair");
    bad += check<BlsFq>("BLS12-381 Fq2 lane pair");
    printf(bad ? "PAIR CHECK FAILED\n" : "PAIR CHECK PASSED\n");
    return bad ? 1 : 0;
}
