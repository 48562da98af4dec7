#pragma once
#include <cstdint>
namespace sb {
struct MsmGeom {
    int c;            // window bits
    int W;            // windows
    uint32_t B;       // buckets per window = 2^(c-1)
    // precomputed-window mode (registered bases): the table holds 2^(c*w) * P_i at [w*stride + i], every window
    // shares ONE bucket set (key = |digit|-1, value = table index), so there is a single bucket reduction and no
    // Horner recombination.  first = index of this call's point 0 inside the registered set.
    int precomp = 0;
    uint64_t stride = 0, first = 0;
    uint32_t windows() const { return precomp ? 1u : (uint32_t)W; }
    // points the bucket reduction returns per MSM: every window sum comes back as 5 parts with power-of-two weights
    // (msm.cuh k_ws_final); the host applies the weights (a few doublings of single points, microseconds on a CPU core)
    uint32_t wsum_points() const { return 5u * windows(); }
};
}
