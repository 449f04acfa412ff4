// Tile layout of the per-feature matrix K = H_x P H_x^T + sigma^2 I that the front-end kernels hand to each other through a
// global (L2 / HBM) buffer, shared by front_split.cu (tracks of up to 32 observations), front.cu (its geometry-only mode for
// longer tracks) and front_chol_smem.cu.
//
// K is padded to 8x8 tiles and stored as the lower tile triangle, tile (I, J) at tile_index(I, J), every tile in the
// accumulator-fragment layout of the FP64 tensor instruction (DMMA.8x8x4): double2 number 4 g + t holds the elements
// (row g, columns 2 t, 2 t + 1) -- i.e. plain row-major inside the tile.  Rows [0, n2) are the measurement rows (n2 = 2 M),
// [n2, a4) identity (a4 = n2 rounded up to a multiple of 4), [a4, a4 + 4) the appended rows [r~; Q1^T] (they never straddle
// a tile row and their diagonal block is zero: after the factorisation of the columns < a4 it holds minus their Gram
// matrix, from which gamma follows), the rest zero.  At M = 30 the appended rows fill the four padding rows of the eighth
// tile row.
#pragma once
#include "msckf_common.cuh"

namespace ftile {

constexpr int PB = 38;                         // doubles per staged 6x6 block of P (stride 76 words = 12 mod 32)
constexpr int AR = 14;                         // doubles per observation's 2x6 block in shared memory (28 words)

__host__ __device__ __forceinline__ int tile_index(int I, int J) { return I * (I + 1) / 2 + J; }

// tile rows of the padded matrix of a track of M observations, and tiles of the lower triangle for an observation capacity
__host__ __device__ __forceinline__ int tile_rows(int M) { return ((((2 * M + 3) & ~3) + 4 + 7) >> 3); }
__host__ __device__ __forceinline__ int tiles_for_capacity(int Mcap) { const int T = tile_rows(Mcap); return T * (T + 1) / 2; }

__device__ __forceinline__ void store_block(double2* Kt, const int i, const int k, const double o00, const double o01,
                                            const double o10, const double o11) {
    // rows 2 i, 2 i + 1 / columns 2 k, 2 k + 1  ->  tile (i / 4, k / 4), lanes 4 (2 (i & 3) + row) + (k & 3)
    const int I = i >> 2, J = k >> 2;
    double2* T = Kt + tile_index(I, J) * 32;
    const int l0 = 8 * (i & 3) + (k & 3);
    T[l0] = make_double2(o00, o01);
    T[l0 + 4] = make_double2(o10, o11);
    if (I == J && i != k) {   // the mirrored block of a diagonal tile is never read as data: keep it defined
        const int l1 = 8 * (k & 3) + (i & 3);
        T[l1] = make_double2(0.0, 0.0);
        T[l1 + 4] = make_double2(0.0, 0.0);
    }
}

// pair p = i (i + 1) / 2 + k of the flat lower-triangle list
__device__ __forceinline__ void pair_of(const int p, int& i, int& k) {
    i = (int)((sqrtf(8.0f * (float)p + 1.0f) - 1.0f) * 0.5f);
    i -= (i * (i + 1) / 2 > p) ? 1 : 0;
    i += ((i + 1) * (i + 2) / 2 <= p) ? 1 : 0;
    k = p - i * (i + 1) / 2;
}

// columns (2 ob, 2 ob + 1) of the four appended rows [r~; Q1^T] for observation ob of a track of M observations
// (hf = the observation's two rows of Q1: hf[0..2] row 0, hf[3..5] row 1)
__device__ __forceinline__ void write_appended_obs(double2* Kt, const int M, const int ob, const double rt0, const double rt1,
                                                   const double h00, const double h01, const double h02, const double h10,
                                                   const double h11, const double h12) {
    const int n2 = 2 * M, a4 = (n2 + 3) & ~3;
    const int Tn = (a4 + 4 + 7) >> 3;
    const int rr = a4 - 8 * (Tn - 1);          // local row of the first appended row in the last tile row: 0 or 4
    const int ia = a4 >> 1;                     // the appended rows as two observation-sized row pairs ia, ia + 1
    double2* Gt = Kt + tile_index(Tn - 1, ob >> 2) * 32 + 4 * rr + (ob & 3);
    Gt[0] = make_double2(rt0, rt1);
    Gt[4] = make_double2(h00, h10);
    Gt[8] = make_double2(h01, h11);
    Gt[12] = make_double2(h02, h12);
    if ((ob >> 2) == Tn - 1) {   // measurement rows of the same (diagonal) tile: define the mirrored blocks
        double2* Tm = Kt + tile_index(Tn - 1, Tn - 1) * 32 + 8 * (ob & 3) + (ia & 3);
        Tm[0] = Tm[1] = Tm[4] = Tm[5] = make_double2(0.0, 0.0);
    }
}

// the padding row pairs i = M .. Mp-1 (at most 7): identity before the appended rows, zero after; the blocks (appended
// rows x measurement columns) are written by write_appended_obs.  `idx` of `stride` cooperating threads.
__device__ __forceinline__ void write_padding_rows(double2* Kt, const int M, const int idx, const int stride) {
    const int n2 = 2 * M, a4 = (n2 + 3) & ~3;
    const int Tn = (a4 + 4 + 7) >> 3, Mp = 4 * Tn;
    const int ia = a4 >> 1;
    for (int i = M; i < Mp; ++i) {
        const bool appended = i >= ia && i < ia + 2;
        for (int k = (appended ? M : 0) + idx; k <= i; k += stride) {
            const double dg = (i == k && i < ia) ? 1.0 : 0.0;
            const int I = i >> 2, J = k >> 2;
            double2* T = Kt + tile_index(I, J) * 32;
            const int l0 = 8 * (i & 3) + (k & 3);
            T[l0] = make_double2(dg, 0.0);
            T[l0 + 4] = make_double2(0.0, dg);
            if (I == J && i != k) {
                const int l1 = 8 * (k & 3) + (i & 3);
                T[l1] = make_double2(0.0, 0.0);
                T[l1 + 4] = make_double2(0.0, 0.0);
            }
        }
    }
}

}  // namespace ftile
