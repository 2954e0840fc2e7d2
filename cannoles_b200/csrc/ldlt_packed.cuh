// ldlt_packed.cuh -- CTA-level dense pivot-free LDL^T on an index-mapped column layout, shared by the
// batched kernels (packed triangle of a KKT system) and k_front_small (m x m front in shared memory).
#pragma once
#include "b2_cuda.h"
#include "mma.cuh"

namespace b2 {

// ------------------------------------------------------------------------------------------
// Dense pivot-free LDL^T of nb <= 64 pivot columns of a supernode panel held in the PACKED layout,
// right-looking over the m rows and the wl >= nb columns of the panel (the rows beyond the pivot block
// are substituted, the columns nb .. wl - 1 of the panel are updated; what lies to the right of the
// panel belongs to other supernodes, which fetch it left-looking): the CTA-level scheme of the single-system engine (cta_ldlt64 in
// kernels.cuh) on the accessor  S(i, j) = Pk[cbm[c0 + o + j] + rowmap[o + i]],  i >= j.
//   warp 0 owns the pivot chain: every lane factors the 8 x 8 diagonal sub-block redundantly in
//   registers, TWO pivots per step (2 x 2 leading block [a b; b c]: x_ij -= (p_i u_j + q_i v_j) / det,
//   p = c u - b v, q = a v - b u, det = a c - b^2; D stays diagonal: d_k = a, d_{k+1} = det / a);
//   (B) rows below the 8 x 8 block: substitution, one row per thread (L to S, W = L D to Wd);
//   (C) warps 1.. apply the rank-8 update to the trailing block while warp 0 updates only the next
//       8 x 8 diagonal block and goes straight on to factor it (look-ahead inside the CTA).
// Wd: 8 m + 8 doubles of shared memory.  On return strict lower = L, diagonal = D.
// (All index lookups are hoisted into registers by hand: a store to Pk may alias cbm / rowmap as far
// as the compiler knows, and a reload of the index before every access doubles the latency of
// loops that are latency-bound -- 66 k instead of 51 k cycles for the 65-column root, measured.)
struct PackedAcc {
  double* Pk;
  const int32_t* cbm;      // shared-memory copy of the column bases
  const int32_t* rowmap;   // local row of the supernode -> permuted index
  int c0, o;               // first column of the supernode, first column of this block inside it
  __device__ __forceinline__ int colbase(int j) const { return cbm[c0 + o + j]; }
  __device__ __forceinline__ int rowidx(int i, int wl) const { return i < wl ? c0 + o + i : rowmap[o + i]; }
};

// warp 0: the 8 x 8 block whose column bases are cb[0..7] and whose first row is r0 (pw <= 8 valid columns)
static __device__ __noinline__ void warp_ldlt8_packed(double* Pk, const int32_t* cbp, int r0, int pw, double* rds) {
  double g[8][8], rd[8];
  int cb[8];
  B2_UNROLL
  for (int t = 0; t < 8; t++) cb[t] = t < pw ? cbp[t] : 0;
  B2_UNROLL
  for (int c = 0; c < 8; c++)
    B2_UNROLL
    for (int t = 0; t <= c; t++) g[c][t] = (c < pw) ? Pk[cb[t] + r0 + c] : (c == t ? 1.0 : 0.0);
  B2_UNROLL
  for (int k = 0; k < 8; k += 2) {
    const double a = g[k][k], b = g[k + 1][k], c = g[k + 1][k + 1];
    const double det = __fma_rn(a, c, -(b * b));
    const double rdet = rcp_nr(det);
    const double ra = rcp_nr(a);
    double p[8], q[8];
    B2_UNROLL
    for (int i = k + 2; i < 8; i++) {
      p[i] = __fma_rn(c, g[i][k], -(b * g[i][k + 1]));
      q[i] = __fma_rn(a, g[i][k + 1], -(b * g[i][k]));
    }
    B2_UNROLL
    for (int i = k + 2; i < 8; i++)
      B2_UNROLL
      for (int j = k + 2; j <= i; j++) {
        const double num = __fma_rn(p[i], g[j][k], q[i] * g[j][k + 1]);
        g[i][j] = __fma_rn(-num, rdet, g[i][j]);
      }
    B2_UNROLL
    for (int i = k + 2; i < 8; i++) { g[i][k] *= ra; g[i][k + 1] = q[i] * rdet; }
    g[k + 1][k] = b * ra;
    g[k + 1][k + 1] = det * ra;
    rd[k] = ra;
    rd[k + 1] = a * rdet;
  }
  __syncwarp();   // every lane has read the unfactored block before anyone overwrites it
  B2_UNROLL
  for (int c = 0; c < 8; c++) {
    if (c < pw) {
      B2_UNROLL
      for (int t = 0; t <= c; t++) Pk[cb[t] + r0 + c] = g[c][t];
      rds[c] = rd[c];
    }
  }
}

template <int NT>
__device__ __forceinline__ void cta_ldlt_packed(PackedAcc S, int nb, int wl, int m, double* Wd, double* rds) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* const Pk = S.Pk;
  const int g0 = S.c0 + S.o;            // permuted index of local row / column 0 of this block
  if (warp == 0) warp_ldlt8_packed(Pk, S.cbm + g0, g0, min(8, nb), rds);
  for (int kb = 0; kb < nb; kb += 8) {
    const int pw = min(8, nb - kb);
    const int t0 = kb + 8;
    __syncthreads();                    // L8 / 1/d of this panel are in S / rds; the trailing block is up to date
    int cb8[8];
    B2_UNROLL
    for (int c = 0; c < 8; c++) cb8[c] = c < pw ? S.colbase(kb + c) : 0;
    // (B) rows below the 8 x 8 block: w[c] = a[c] - sum_{t<c} w[t] L8[c][t],  l[c] = w[c] / d_c
    // (a partial last panel, pw < 8, is followed directly by the rows below the pivot block)
    for (int row = kb + pw + tid; row < m; row += NT) {
      const int rr = S.rowidx(row, wl);
      double wv[8], l8[8][8], rdv[8];
      B2_UNROLL
      for (int c = 0; c < 8; c++) {
        wv[c] = (c < pw) ? Pk[cb8[c] + rr] : 0.0;
        rdv[c] = c < pw ? rds[c] : 0.0;
      }
      B2_UNROLL
      for (int c = 1; c < 8; c++)
        B2_UNROLL
        for (int t = 0; t < c; t++) l8[c][t] = (c < pw) ? Pk[cb8[t] + g0 + kb + c] : 0.0;
      B2_UNROLL
      for (int c = 1; c < 8; c++)
        B2_UNROLL
        for (int t = 0; t < c; t++) wv[c] -= wv[t] * l8[c][t];
      B2_UNROLL
      for (int c = 0; c < 8; c++) {
        if (c < pw) Pk[cb8[c] + rr] = wv[c] * rdv[c];
        Wd[row * 8 + c] = c < pw ? wv[c] : 0.0;
      }
    }
    __syncthreads();                    // panel kb complete
    if (t0 >= wl) break;                // no column of the panel left to update
    if (warp == 0) {
      if (t0 < nb) {
        // the next 8 x 8 diagonal block only: lane <-> (row i, columns 2 jp, 2 jp + 1), then its LDL^T
        const int i = t0 + (lane >> 2), j = t0 + 2 * (lane & 3);
        if (i < nb) {
          const int ri = g0 + i;
          const int cbj = S.colbase(j), cbj1 = j + 1 < wl ? S.colbase(j + 1) : 0;
          double lv[8];
          B2_UNROLL
          for (int c = 0; c < 8; c++) lv[c] = Pk[cb8[c] + ri];
          double s0 = 0.0, s1 = 0.0;
          B2_UNROLL
          for (int c = 0; c < 8; c++) {
            if (j < m) s0 += lv[c] * Wd[j * 8 + c];
            if (j + 1 < m) s1 += lv[c] * Wd[(j + 1) * 8 + c];
          }
          if (j <= i) Pk[cbj + ri] -= s0;
          if (j + 1 <= i) Pk[cbj1 + ri] -= s1;
        }
        __syncwarp();
        warp_ldlt8_packed(Pk, S.cbm + g0 + t0, g0 + t0, min(8, nb - t0), rds);
      }
    } else {
      // (C) rows t1 .. m - 1 (warp 0 has the pivot rows t0 .. min(t0 + 8, nb) - 1), columns t0 .. row.
      // Thread <-> (row pair, column phase): rows t1 + p and m - 1 - p together have the same number
      // of columns whatever p; the phases (one per warp) take every NPH-th column.
      const int t1 = t0 < nb ? min(t0 + 8, nb) : t0;
      const int n = m - t1;
      const int phase = warp - 1;
      constexpr int NPH = NT / 32 - 1;
      for (int p = lane; p < (n + 1) / 2; p += 32) {
        const int iA = t1 + p, iB = m - 1 - p;
        const bool two = iB > iA;
        const int rA = S.rowidx(iA, wl), rB = S.rowidx(iB, wl);
        double la[8], lb[8];
        B2_UNROLL
        for (int c = 0; c < 8; c++) {
          la[c] = c < pw ? Pk[cb8[c] + rA] : 0.0;
          lb[c] = (two && c < pw) ? Pk[cb8[c] + rB] : 0.0;
        }
        const int jmax = min(two ? iB : iA, wl - 1);
#pragma unroll 2
        for (int j = t0 + phase; j <= jmax; j += NPH) {
          const double* wj = Wd + j * 8;
          const int cbj = S.colbase(j);
          double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
          B2_UNROLL
          for (int c = 0; c < 8; c += 2) {
            const double w0 = wj[c], w1 = wj[c + 1];
            a0 += la[c] * w0; a1 += la[c + 1] * w1;
            b0 += lb[c] * w0; b1 += lb[c + 1] * w1;
          }
          if (j <= iA) Pk[cbj + rA] -= a0 + a1;
          if (two) Pk[cbj + rB] -= b0 + b1;
        }
      }
    }
  }
}

}  // namespace b2
