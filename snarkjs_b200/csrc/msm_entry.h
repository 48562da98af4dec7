// msm_entry.h — untyped per-(curve, group) MSM entry points; each group/curve pair lives in its own translation
// unit so the four heavy template instantiations compile in parallel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "msm_geom.h"
namespace sb {
struct MsmScratch; struct MsmLaunchStats; struct MsmSorted;
// coordinate bytes (n8q or 2*n8q) of the group's field
#define SB_DECL_GROUP(NAME) \
    /* async: bucket pipeline for one base set over sorted entries; d_wsum = device buffer of W XYZZ points */ \
    int NAME##_buckets(const void* d_bases, const MsmSorted& s, MsmScratch& scratch, cudaStream_t stream, void* d_wsum, MsmLaunchStats* stats, \
                       cudaStream_t tail_stream, cudaEvent_t ev_acc); \
    /* host: Horner over W window sums (host bytes) added into acc_xyzz (host XYZZ bytes, in/out) */ \
    void NAME##_combine(const uint8_t* wsum_host, const MsmGeom& g, uint8_t* acc_xyzz); \
    /* host: acc_xyzz += other */ \
    void NAME##_add(uint8_t* acc_xyzz, const uint8_t* other_xyzz); \
    /* host: XYZZ -> normalised Jacobian bytes (3 coordinates) */ \
    void NAME##_to_jacobian(const uint8_t* xyzz, uint8_t* out); \
    /* host: XYZZ -> affine bytes (2 coordinates, infinity = zeros) */ \
    void NAME##_to_affine(const uint8_t* xyzz, uint8_t* out); \
    /* host: affine bytes -> XYZZ */ \
    void NAME##_from_affine(const uint8_t* aff, uint8_t* xyzz); \
    /* host: out = k * p, k plain little-endian scalar of nbytes */ \
    void NAME##_times(const uint8_t* xyzz, const uint8_t* k, int nbytes, uint8_t* out); \
    /* async: n synthetic points (k0(c) + j*kd)*G, see msm.cuh k_gen_points into d_out; gen = affine generator bytes (host) */ \
    int NAME##_gen_points(const uint8_t* gen_affine, uint64_t seed, uint64_t n, void* d_out, cudaStream_t stream); \
    /* async: table[w*n + i] = 2^(c*w) * bases[i], w < W (device pointers) */ \
    int NAME##_precompute(const void* d_bases, uint64_t n, int c, int W, void* d_table, cudaStream_t stream); \
    uint32_t NAME##_xyzz_bytes();
SB_DECL_GROUP(bn254_g1)
SB_DECL_GROUP(bn254_g2)
SB_DECL_GROUP(bls12381_g1)
SB_DECL_GROUP(bls12381_g2)
}
