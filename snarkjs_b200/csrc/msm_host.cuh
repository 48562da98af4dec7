// msm_host.cuh — host side of one MSM: window recombination and normalisation.
// Mirrors the recombination in _multiExpChunk (reference build/snarkjs.js:14594-14600: from the top window
// down, c doublings then add) and the chunk sum in _multiExp (14655-14658).
#pragma once
#include <vector>
#include <cstring>
#include "msm.cuh"

namespace sb {

template <class F> void msm_combine_host(const XYZZ<F>* ws, const MsmGeom& g, XYZZ<F>& total) {
    XYZZ<F> r = XYZZ<F>::inf();
    for (int w = g.W - 1; w >= 0; w--) {
        if (!r.is_inf()) for (int j = 0; j < g.c; j++) r = XYZZ<F>::dbl(r);
        r.add(ws[w]);
    }
    total.add(r);
}

// (x, y, 1) Montgomery Jacobian bytes; infinity -> (0, 1, 0) (reference 6039-6065)
template <class F> void xyzz_to_jacobian_bytes(const XYZZ<F>& p, uint8_t* out) {
    const size_t n8 = sizeof(F);
    F x, y, z;
    if (p.is_inf()) { x = F::zero(); y = F::one(); z = F::zero(); }
    else { x = F::mul(p.x, F::inv(p.zz)); y = F::mul(p.y, F::inv(p.zzz)); z = F::one(); }
    memcpy(out, &x, n8); memcpy(out + n8, &y, n8); memcpy(out + 2 * n8, &z, n8);
}

}  // namespace sb
