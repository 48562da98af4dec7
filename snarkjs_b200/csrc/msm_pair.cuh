// msm_pair.cuh — batched-affine pre-reduction of the sorted MSM entry list.
//
// The bucket accumulation of msm.cuh adds points one by one into an XYZZ accumulator: 10 modmul per term.  In affine
// coordinates an addition costs 3 modmul + one field inversion; with Montgomery's simultaneous-inversion trick the
// inversion is shared by a whole batch (3 more modmul per addition), i.e. ~6 modmul per term — if thousands of
// *independent* additions are available at once.  They are: inside every bucket the points can be summed pairwise,
// (p0+p1), (p2+p3), ..., which halves the bucket, and all pairs of all buckets are independent.
//
//   k_bucket_offsets   off[b] = first sorted entry of bucket b (binary search on the sorted keys)
//   per round r (R rounds, R ~ log2(average bucket size) + 1):
//     k_halve_sizes + cub::DeviceScan   off_out = exclusive_scan(ceil(size_in / 2))
//     k_pair_round       output slot o of bucket b, j = o - off_out[b]:  dst[o] = src[off_in[b]+2j] + src[off_in[b]+2j+1]
//                        (or a copy when the bucket size is odd and this is its last element).  One CTA (128 threads x K
//                        slots) shares ONE field inversion: per-thread prefix products, a shared-memory scan of the 128
//                        thread products, Fermat inversion of the CTA total by warp 0, back-substitution.
//   The list that remains (<= 1 point per bucket for uniform scalars; longer runs survive only for giant buckets)
//   goes through the unchanged segmented XYZZ accumulation / fold / reduce of msm.cuh, which keeps every special case.
//
// Special cases inside a pair (all exercised by tests/test_gpu_parity.py::test_msm_edge_cases): either point at infinity
// ((0,0), reference build/snarkjs.js:6068-6086), P + P (doubling: denominator 2y, numerator 3x^2), P + (-P) (infinity).
#pragma once
#include <cuda_runtime.h>
#include "ec.cuh"
#include "msm_geom.h"

namespace sb {

static constexpr int PAIR_K = 32;          // output slots per thread
static constexpr int PAIR_THREADS = 128;

// off[b] = lower_bound(keys[0..M), b) for b in [0, NB]; M = counts[0]
__global__ void k_bucket_offsets(const uint32_t* __restrict__ keys, const uint64_t* __restrict__ counts, uint32_t NB, uint32_t* __restrict__ off);
// sizes[b] = ceil((off[b+1]-off[b]) / 2), sizes[NB] = 0
__global__ void k_halve_sizes(const uint32_t* __restrict__ off, uint32_t NB, uint32_t* __restrict__ sizes);
// counts[0] = off[NB] and the fold-cascade level sizes derived from it
__global__ void k_counts_from_offsets(const uint32_t* __restrict__ off, uint32_t NB, uint64_t* __restrict__ counts);

template <class F> struct PairInv;   // field inversion usable on the device (Fp: Fermat; Fp2: norm)
template <class P> struct PairInv<Fp<P>> { __device__ static Fp<P> inv(const Fp<P>& a) { return Fp<P>::inv_binary(a); } };
template <class P> struct PairInv<Fp2<P>> { __device__ static Fp2<P> inv(const Fp2<P>& a) { return Fp2<P>::inv(a); } };

template <class F> __device__ __forceinline__ F pair_load_x(const Affine<F>* p) { return load_vec(&p->x); }
template <class F> __device__ __forceinline__ F pair_load_y(const Affine<F>* p) { return load_vec(&p->y); }

// classification of one output slot; den is the value that enters the shared inversion
enum { PAIR_COPY = 0, PAIR_ADD = 1, PAIR_DBL = 2, PAIR_INF = 3, PAIR_TAKE2 = 4, PAIR_NONE = 5 };

template <class F, bool FIRST>
__device__ __forceinline__ const Affine<F>* pair_src(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ vals, uint32_t s, bool& neg) {
    if (FIRST) { uint32_t v = vals[s]; neg = (v >> 31) != 0; return pts + (v & 0x7fffffffu); }
    neg = false; return pts + s;
}

// Decides what slot (s0, pair?) does and returns the denominator (Montgomery one when no inversion is needed).
template <class F, bool FIRST>
__device__ __forceinline__ int pair_classify(const Affine<F>* __restrict__ pts, const uint32_t* __restrict__ vals, uint32_t s0, bool has_pair, F& den) {
    den = F::one();
    if (!has_pair) return PAIR_COPY;
    bool n1, n2;
    const Affine<F>* p1 = pair_src<F, FIRST>(pts, vals, s0, n1);
    const Affine<F>* p2 = pair_src<F, FIRST>(pts, vals, s0 + 1, n2);
    F x1 = pair_load_x<F>(p1), x2 = pair_load_x<F>(p2);
    if (x1.is_zero()) { if (pair_load_y<F>(p1).is_zero()) return PAIR_TAKE2; }      // p1 = infinity
    if (x2.is_zero()) { if (pair_load_y<F>(p2).is_zero()) return PAIR_COPY; }       // p2 = infinity
    if (x1 == x2) {
        F y1 = F::cneg(pair_load_y<F>(p1), n1), y2 = F::cneg(pair_load_y<F>(p2), n2);
        if (y1 == y2 && !y1.is_zero()) { den = F::dbl(y1); return PAIR_DBL; }
        return PAIR_INF;
    }
    den = F::sub(x2, x1);
    return PAIR_ADD;
}

template <class F, bool FIRST>
__global__ void __launch_bounds__(PAIR_THREADS)
k_pair_round(const Affine<F>* __restrict__ src, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ off_in,
             const uint32_t* __restrict__ off_out, uint32_t NB, Affine<F>* __restrict__ dst, uint32_t* __restrict__ dst_keys,
             F* __restrict__ scratch) {
    __shared__ F sP[PAIR_THREADS];   // prefix products of the thread totals
    __shared__ F sS[PAIR_THREADS];   // suffix products
    __shared__ F sInv;
    const uint32_t total_out = off_out[NB];
    const uint32_t tid = threadIdx.x;
    const uint64_t gt = blockIdx.x * (uint64_t)blockDim.x + tid;
    const uint64_t nthreads = gridDim.x * (uint64_t)blockDim.x;
    const uint64_t o0 = gt * PAIR_K;
    if ((uint64_t)blockIdx.x * blockDim.x * PAIR_K >= total_out) return;        // whole CTA idle
    // bucket of slot o0: last b with off_out[b] <= o0 (upper_bound - 1)
    uint32_t b0 = 0;
    if (o0 < total_out) {
        uint32_t lo = 0, hi = NB;                                               // invariant: off_out[lo] <= o0 < off_out[hi]
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (off_out[mid] <= o0) lo = mid; else hi = mid; }
        b0 = lo;
    }
    // ---- pass A: prefix products of the denominators
    F pre = F::one();
    {
        uint32_t b = b0;
        for (int k = 0; k < PAIR_K; k++) {
            uint64_t o = o0 + k;
            F den = F::one();
            if (o < total_out) {
                while (off_out[b + 1] <= o) b++;
                uint32_t j = (uint32_t)o - off_out[b], s0 = off_in[b] + 2 * j, m = off_in[b + 1] - off_in[b];
                pair_classify<F, FIRST>(src, vals, s0, 2 * j + 1 < m, den);
            }
            store_vec(scratch + (uint64_t)k * nthreads + gt, pre);
            pre = F::mul(pre, den);
        }
    }
    // ---- one inversion per CTA: inverse of thread t's product = inv(total) * prefix(t-1) * suffix(t+1)
    store_vec(&sP[tid], pre); store_vec(&sS[tid], pre);
    __syncthreads();
    for (int d = 1; d < PAIR_THREADS; d <<= 1) {
        F a, c; bool ha = tid >= (uint32_t)d, hc = tid + d < PAIR_THREADS;
        if (ha) a = F::mul(load_vec(&sP[tid - d]), load_vec(&sP[tid]));
        if (hc) c = F::mul(load_vec(&sS[tid]), load_vec(&sS[tid + d]));
        __syncthreads();
        if (ha) store_vec(&sP[tid], a);
        if (hc) store_vec(&sS[tid], c);
        __syncthreads();
    }
    if (tid < 32) {                                    // warp 0 computes the inverse (all lanes: same cost as one)
        F t = PairInv<F>::inv(load_vec(&sP[PAIR_THREADS - 1]));
        if (tid == 0) store_vec(&sInv, t);
    }
    __syncthreads();
    F running = load_vec(&sInv);
    if (tid > 0) running = F::mul(running, load_vec(&sP[tid - 1]));
    if (tid + 1 < PAIR_THREADS) running = F::mul(running, load_vec(&sS[tid + 1]));
    // ---- pass B: back-substitution, last slot first
    if (o0 >= total_out) return;
    uint32_t b = b0;
    {   // bucket of the last active slot
        uint64_t olast = o0 + PAIR_K - 1; if (olast >= total_out) olast = total_out - 1;
        while (off_out[b + 1] <= olast) b++;
    }
    for (int k = PAIR_K - 1; k >= 0; k--) {
        uint64_t o = o0 + k;
        if (o >= total_out) continue;
        while (off_out[b] > o) b--;
        uint32_t j = (uint32_t)o - off_out[b], s0 = off_in[b] + 2 * j, m = off_in[b + 1] - off_in[b];
        F den;
        int kind = pair_classify<F, FIRST>(src, vals, s0, 2 * j + 1 < m, den);
        Affine<F> out;
        bool n1, n2;
        if (kind == PAIR_ADD || kind == PAIR_DBL) {
            F inv = F::mul(running, load_vec(scratch + (uint64_t)k * nthreads + gt));
            running = F::mul(running, den);
            const Affine<F>* p1 = pair_src<F, FIRST>(src, vals, s0, n1);
            const Affine<F>* p2 = pair_src<F, FIRST>(src, vals, s0 + 1, n2);
            F x1 = pair_load_x<F>(p1), y1 = F::cneg(pair_load_y<F>(p1), n1);
            F x2 = pair_load_x<F>(p2);
            F lam;
            if (kind == PAIR_ADD) { F y2 = F::cneg(pair_load_y<F>(p2), n2); lam = F::mul(F::sub(y2, y1), inv); }
            else { F xx = F::sqr(x1); lam = F::mul(F::add(F::dbl(xx), xx), inv); }
            F x3 = F::sub(F::sub(F::sqr(lam), x1), x2);
            out.x = x3; out.y = F::sub(F::mul(lam, F::sub(x1, x3)), y1);
        } else if (kind == PAIR_INF) {
            out.x = F::zero(); out.y = F::zero();
        } else {                                       // PAIR_COPY (single or p2 = inf) / PAIR_TAKE2 (p1 = inf)
            const Affine<F>* p = pair_src<F, FIRST>(src, vals, kind == PAIR_TAKE2 ? s0 + 1 : s0, n1);
            out.x = pair_load_x<F>(p); out.y = F::cneg(pair_load_y<F>(p), n1);
        }
        store_vec(dst + o, out);
        dst_keys[o] = b;
    }
}

}  // namespace sb
