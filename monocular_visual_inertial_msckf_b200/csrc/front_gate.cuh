// The end of the per-feature factorisation, shared by the register-resident kernel (front_split.cu) and the shared-memory
// kernel (front_chol_smem.cu): the last diagonal tile of the padded K (front_tiles.cuh) in accumulator-fragment layout
// (lane (g, t): D0 = element (g, 2 t), D1 = (g, 2 t + 1)), already updated by every tile column to its left.
#pragma once
#include "front_geom.cuh"

namespace ftile {

using fgeom::fast_rsqrt;
using fgeom::fast_rsqrt_spd;
using fgeom::shfl_d;

// the first `npass` column pairs of a diagonal tile, no inverse (last tile row: the columns left of the appended rows)
__device__ __forceinline__ void diag_tile_partial(double& D0, double& D1, const int g, const int t, const int npass) {
#pragma unroll 1
    for (int tk = 0; tk < npass; ++tk) {
        {
            const int k = 2 * tk;
            const double rs = fast_rsqrt_spd(shfl_d(D0, k * 4 + tk));
            const double colv = (g > k) ? D0 * rs : 0.0;
            const double lg = shfl_d(colv, g * 4 + tk);
            const double lc0 = shfl_d(colv, (2 * t) * 4 + tk);
            const double lc1 = shfl_d(colv, (2 * t + 1) * 4 + tk);
            D0 -= lg * lc0;
            D1 -= lg * lc1;
        }
        {
            const int k = 2 * tk + 1;
            const double rs = fast_rsqrt_spd(shfl_d(D1, k * 4 + tk));
            const double colv = (g > k) ? D1 * rs : 0.0;
            const double lg = shfl_d(colv, g * 4 + tk);
            const double lc0 = shfl_d(colv, (2 * t) * 4 + tk);
            const double lc1 = shfl_d(colv, (2 * t + 1) * 4 + tk);
            D0 -= lg * lc0;
            D1 -= lg * lc1;
        }
    }
}


// gamma = y.y - (V^T y)^T (V^T V)^-1 (V^T y) from the last diagonal tile: its first rr columns are measurement / identity
// columns still to be factored; then the 4x4 block at (rr, rr) holds MINUS the Gram matrix of [y; V] = [r~; Q1^T] L^-T:
// yy = (0,0), b_q = (q,0), M_qp = (q,p), q >= p (lower triangle of the tile).  rank = rank(H_f).  Whole warp.
__device__ __forceinline__ double gamma_from_last_tile(double D0, double D1, const int g, const int t, const int rr, const int rank) {
    diag_tile_partial(D0, D1, g, t, rr >> 1);
    const int l0 = 4 * rr + (rr >> 1);                  // lane of element (rr, rr)
    const double yy = -shfl_d(D0, l0);
    double bq[3], Mq[6];
    bq[0] = -shfl_d(D0, l0 + 4); bq[1] = -shfl_d(D0, l0 + 8); bq[2] = -shfl_d(D0, l0 + 12);
    Mq[0] = -shfl_d(D1, l0 + 4);                               // (1,1)
    Mq[1] = -shfl_d(D1, l0 + 8);                               // (2,1)
    Mq[2] = -shfl_d(D1, l0 + 12);                              // (3,1)
    Mq[3] = -shfl_d(D0, l0 + 9);                               // (2,2)
    Mq[4] = -shfl_d(D0, l0 + 13);                              // (3,2)
    Mq[5] = -shfl_d(D1, l0 + 13);                              // (3,3)
    // Cholesky solve on the leading rank x rank block of V V^T (reciprocal square roots: 0 for a non-positive pivot)
    const double r0 = (rank >= 1) ? fast_rsqrt(Mq[0]) : 0.0;
    const double z0 = bq[0] * r0;
    const double l10 = Mq[1] * r0, l20 = Mq[2] * r0;
    const double r1 = (rank >= 2 && r0 > 0.0) ? fast_rsqrt(Mq[3] - l10 * l10) : 0.0;
    const double z1 = (bq[1] - l10 * z0) * r1;
    const double l21 = (Mq[4] - l20 * l10) * r1;
    const double r2 = (rank >= 3 && r1 > 0.0) ? fast_rsqrt(Mq[5] - l20 * l20 - l21 * l21) : 0.0;
    const double z2 = (bq[2] - l20 * z0 - l21 * z1) * r2;
    return yy - (z0 * z0 + z1 * z1 + z2 * z2);
}

}  // namespace ftile
