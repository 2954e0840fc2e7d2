// kernels.cuh -- sm_100a device code of the single-system multifrontal LDL^T engine.
//
// Kernel (1)  k_assemble_csc      COO -> CSC accumulate   (set_vals!, src/solver_types.jl:53-59)
// Kernel (2)  k_front_small       one CTA per front, front resident in shared memory
//             k_assemble_large / k_diag_factor / k_trsm / k_update   tiled path for big fronts
//             k_inertia           pivot-sign counts        (src/solver_types.jl:90-96)
// Kernel (3)  k_fwd / k_bwd       level-scheduled supernodal triangular solves
//             k_spmv_sym          y = K x for residuals / iterative refinement
//
// Data layout in HBM (all column-major, Float64):
//   nzval[nnzA]           CSC-upper values of K in the reference's slot order
//   Lx[lptr[s] + i + j*m] panel of front s: m rows (w pivots + r rows below) x w pivot columns;
//                         strict lower part = L, diagonal = D, strict upper of the pivot block = 0
//   CB[cbptr[s] + i + j*r] contribution block (Schur complement) of front s, lower triangle valid
//   dvec[N]               pivots in elimination order
#pragma once
#include "b2_cuda.h"
#include "mma.cuh"
#include "plan.h"

namespace b2 {

// ------------------------------------------------------------------------------------------
// (1) COO -> CSC.  One thread per CSC slot; the duplicates of a slot are summed in increasing
// COO index starting from +0.0, i.e. bit-for-bit the sums `set_vals!` forms.
// Algorithmic bytes: 8 nnz (vals) + 4 nnz (coo_sorted) + 4 nnzA (slot_ptr) + 8 nnzA (nzval).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_assemble_csc(int64_t nslots, const int32_t* __restrict__ slot_ptr,
                                                      const int32_t* __restrict__ coo_sorted,
                                                      const double* __restrict__ vals,
                                                      double* __restrict__ nzval) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  int32_t a = slot_ptr[s], b = slot_ptr[s + 1];
  double acc = 0.0;
  for (int32_t q = a; q < b; q++) acc += vals[coo_sorted[q]];
  nzval[s] = acc;
}

// partial sums "all duplicates but the last" of the rho / delta diagonal slots: with them a
// regularisation retry rewrites nzval[slot] = base + shift, bit-identical to a re-assembly.
__global__ void __launch_bounds__(256) k_diag_base(int n, const int32_t* __restrict__ slots,
                                                   const int32_t* __restrict__ slot_ptr,
                                                   const int32_t* __restrict__ coo_sorted,
                                                   const double* __restrict__ vals,
                                                   double* __restrict__ base) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t s = slots[i];
  int32_t a = slot_ptr[s], b = slot_ptr[s + 1] - 1;
  double acc = 0.0;
  for (int32_t q = a; q < b; q++) acc += vals[coo_sorted[q]];
  base[i] = acc;
}

__global__ void __launch_bounds__(256) k_diag_shift(int n, const int32_t* __restrict__ slots,
                                                    const double* __restrict__ base, double shift,
                                                    double* __restrict__ nzval) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  nzval[slots[i]] = base[i] + shift;
}

// ------------------------------------------------------------------------------------------
// (2a) small fronts: the whole m x m front lives in shared memory.
// zero -> scatter A -> extend-add children (fixed order, deterministic) -> eliminate w pivots
// (right-looking, no pivoting) -> write panel, pivots and contribution block.
// ------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k_front_small(PlanDev P, const int32_t* __restrict__ list, int count) {
  const int b = blockIdx.x;
  if (b >= count) return;
  const int s = list[b];
  B2_DYN_SMEM(raw);
  double* F = reinterpret_cast<double*>(raw);
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int64_t r0 = P.rptr[s];
  const int m = (int)(P.rptr[s + 1] - r0);
  const int r = m - w;
  double* lk = F + (size_t)m * m;
  double* ak = lk + m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;

  for (int idx = tid; idx < m * m; idx += NT) F[idx] = 0.0;
  __syncthreads();
  for (int64_t q = P.amap_ptr[s] + tid; q < P.amap_ptr[s + 1]; q += NT)
    F[P.amap_pos[q]] = P.nzval[P.amap_slot[q]];
  __syncthreads();
  for (int ci = P.child_ptr[s]; ci < P.child_ptr[s + 1]; ci++) {
    const int c = P.child_idx[ci];
    const int wc = P.scol[c + 1] - P.scol[c];
    const int64_t rc0 = P.rptr[c] + wc;
    const int rc = (int)(P.rptr[c + 1] - rc0);
    const int32_t* relc = P.rel + rc0;
    const double* cb = P.CB + P.cbptr[c];
    for (int j = warp; j < rc; j += NW) {
      const int J = relc[j];
      for (int i = j + lane; i < rc; i += 32) F[relc[i] + J * m] += cb[i + (size_t)j * rc];
    }
    __syncthreads();
  }
  for (int k = 0; k < w; k++) {
    const double dk = F[k + k * m];
    for (int i = k + 1 + tid; i < m; i += NT) {
      const double a = F[i + k * m];
      const double l = a / dk;
      ak[i] = a;
      lk[i] = l;
      F[i + k * m] = l;
    }
    if (tid == 0) {
      P.dvec[c0 + k] = dk;
      if (dk == 0.0) P.flags[0] = 1;
    }
    __syncthreads();
    for (int j = k + 1 + warp; j < m; j += NW) {
      const double ajk = ak[j];
      for (int i = j + lane; i < m; i += 32) F[i + j * m] -= lk[i] * ajk;
    }
    __syncthreads();
  }
  double* Lp = P.Lx + P.lptr[s];
  for (int idx = tid; idx < m * w; idx += NT) Lp[idx] = F[idx];
  double* cbp = P.CB + P.cbptr[s];
  for (int j = warp; j < r; j += NW)
    for (int i = j + lane; i < r; i += 32) cbp[i + (size_t)j * r] = F[(w + i) + (w + j) * m];
}

// ------------------------------------------------------------------------------------------
// (2b) tiled path for fronts that do not fit in shared memory.  Pivot blocks of NB = TILE = 64
// columns; per block: diagonal 64 x 64 LDL^T (inside the CTA that produced it) -> k_trsm
// (rows below) -> k_update (trailing pivot columns, FP64 tensor-core tiles); one more k_update
// pass with K = w forms the contribution block.
// ------------------------------------------------------------------------------------------

// item = (front, destination column block, destination row chunk, global column-block id).
// The CTA owns the ASM_ROWS x ASM_COLS destination tile in shared memory: zero, scatter the A
// entries of its columns, add the children's contribution blocks one child after the other
// (fixed order => deterministic sums, no atomics), then write the tile once (panel columns j < w
// go to Lx, the others to CB; only rows >= column are produced).  Which children touch a column
// block, and with which of their columns, is precomputed on the host (asm_cptr / asm_ent): a
// front at the top of the tree has hundreds of small children and scanning them all per tile
// was the whole cost of this kernel.
__global__ void __launch_bounds__(256) k_assemble_large(PlanDev P, const int32_t* __restrict__ items, int nitems) {
  const int b = blockIdx.x;
  if (b >= nitems) return;
  const int s = items[4 * b], j0 = items[4 * b + 1] * ASM_COLS, i0 = items[4 * b + 2] * ASM_ROWS;
  const int gcb = items[4 * b + 3];
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int64_t r0 = P.rptr[s];
  const int m = (int)(P.rptr[s + 1] - r0);
  const int r = m - w;
  const int je = min(j0 + ASM_COLS, m), ie = min(i0 + ASM_ROWS, m);
  __shared__ double T[ASM_COLS][ASM_ROWS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < ASM_COLS * ASM_ROWS; idx += 256) (&T[0][0])[idx] = 0.0;
  __syncthreads();
  if (j0 < w) {  // A entries only land in pivot columns; amap is sorted by position
    const int64_t a0 = P.amap_ptr[s], a1 = P.amap_ptr[s + 1];
    const int lo_pos = j0 * m, hi_pos = min(je, w) * m;
    int64_t lo = a0, hi = a1;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (P.amap_pos[mid] < lo_pos) lo = mid + 1; else hi = mid; }
    const int64_t qa = lo;
    hi = a1;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (P.amap_pos[mid] < hi_pos) lo = mid + 1; else hi = mid; }
    const int64_t qb = lo;
    for (int64_t q = qa + tid; q < qb; q += 256) {
      const int pos = P.amap_pos[q];
      const int j = pos / m, i = pos - j * m;
      if (i >= i0 && i < ie) T[j - j0][i - i0] = P.nzval[P.amap_slot[q]];
    }
  }
  __syncthreads();
  const bool one_chunk = m <= ASM_ROWS;
  for (int64_t e = P.asm_cptr[gcb]; e < P.asm_cptr[gcb + 1]; e++) {
    const int c = P.asm_ent[3 * e], ja = P.asm_ent[3 * e + 1], jz = P.asm_ent[3 * e + 2];
    const int wc = P.scol[c + 1] - P.scol[c];
    const int64_t rc0 = P.rptr[c] + wc;
    const int rc = (int)(P.rptr[c + 1] - rc0);
    const int32_t* relc = P.rel + rc0;
    const double* cb = P.CB + P.cbptr[c];
    int ia = ja, iz = rc;   // rows >= column, so the row range starts no earlier than ja
    if (!one_chunk) {
      int lo = ja, hi = rc;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (relc[mid] < i0) lo = mid + 1; else hi = mid; }
      ia = lo;
      hi = rc;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (relc[mid] < ie) lo = mid + 1; else hi = mid; }
      iz = lo;
    }
    if (ia < iz) {   // uniform over the CTA
      for (int j = ja + warp; j < jz; j += 8) {
        double* dst = &T[relc[j] - j0][0] - i0;
        const double* src = cb + (size_t)j * rc;
#pragma unroll 4
        for (int i = max(ia, j) + lane; i < iz; i += 32) dst[relc[i]] += src[i];
      }
      __syncthreads();
    }
  }
  double* Lp = P.Lx + P.lptr[s];
  double* cbp = P.CB + P.cbptr[s];
  for (int j = j0 + warp; j < je; j += 8) {
    const double* src = &T[j - j0][0] - i0;
    if (j < w) {
      for (int i = max(i0, j) + lane; i < ie; i += 32) Lp[i + (size_t)j * m] = src[i];
    } else {
      for (int i = max(i0, j) + lane; i < ie; i += 32) cbp[(i - w) + (size_t)(j - w) * r] = src[i];
    }
  }
}

// CTA-level pivot-free LDL^T of an nb x nb block (nb <= NB = 64) held column-major in shared
// memory (S[i + j * DIAG_LD], lower triangle), blocked in 8-column panels: warp 0 factors the
// panel (row-per-lane in registers, two rows per lane, pivots and multipliers exchanged with
// shuffles), then all NT threads apply the rank-8 update to the trailing columns.  16 barriers
// instead of 128, small loop bodies (this runs cold, once per pivot block, on the critical path).
// dsh: NB doubles of shared memory.  On return the strict lower part holds L, the diagonal D.
template <int NT>
__device__ __forceinline__ void cta_ldlt64(double* S, int nb, double* dsh, int* flags) {
  constexpr int ld = DIAG_LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ti = tid & 63, tj = tid >> 6;
  for (int kb = 0; kb < nb; kb += 8) {
    if (warp == 0) {
      const int pw = min(8, nb - kb);
      const int r0 = kb + lane, r1 = kb + 32 + lane;
      double p0[8], p1[8];
      B2_UNROLL
      for (int c = 0; c < 8; c++) {
        p0[c] = (r0 < nb && c < pw && kb + c <= r0) ? S[r0 + (kb + c) * ld] : 0.0;
        p1[c] = (r1 < nb && c < pw) ? S[r1 + (kb + c) * ld] : 0.0;
      }
      int bad = 0;
      B2_UNROLL
      for (int c = 0; c < 8; c++) {
        if (c < pw) {
          const double dk = __shfl_sync(0xffffffffu, p0[c], c);   // row kb + c lives in lane c
          if (dk == 0.0) bad = 1;
          const double a0 = p0[c], a1 = p1[c];
          const double l0 = a0 / dk, l1 = a1 / dk;
          B2_UNROLL
          for (int cc = c + 1; cc < 8; cc++) {
            const double ajc = __shfl_sync(0xffffffffu, a0, cc);  // unscaled A(kb + cc, kb + c)
            p0[cc] -= l0 * ajc;
            p1[cc] -= l1 * ajc;
          }
          if (lane > c) p0[c] = l0;
          p1[c] = l1;
        }
      }
      B2_UNROLL
      for (int c = 0; c < 8; c++) {
        if (c < pw) {
          if (r0 < nb && kb + c <= r0) S[r0 + (kb + c) * ld] = p0[c];
          if (r1 < nb) S[r1 + (kb + c) * ld] = p1[c];
          if (lane == c) dsh[kb + c] = p0[c];
        }
      }
      if (bad && lane == 0) flags[0] = 1;
    }
    __syncthreads();
    const int t0 = kb + 8;
    if (t0 < nb) {
#pragma unroll 2
      for (int j = t0 + tj; j < nb; j += NT / 64) {
        const int i = j + ti;
        if (i < nb) {
          double acc = 0.0;
          B2_UNROLL
          for (int c = 0; c < 8; c++)
            acc += S[i + (kb + c) * ld] * (dsh[kb + c] * S[j + (kb + c) * ld]);
          S[i + j * ld] -= acc;
        }
      }
    }
    __syncthreads();
  }
}

// item = (front, row chunk).  Every CTA first factors the nb x nb diagonal block (jb, jb) of the
// panel in shared memory (redundantly: a few us of work instead of one more launch on the
// critical path; chunk 0 stores the factored block in the staging area -- NOT in the panel, which
// sibling CTAs may still be reading -- and the pivots in dvec), then forms
// L21 = A21 L11^{-T} D^{-1} for its TRSM_ROWS rows below the block: one row per thread, the row
// staged in shared memory, substitution in 8-column blocks (rolled loops, 8 x 8 unrolled bodies).
// Dynamic shared memory: TRSM_SMEM bytes.
constexpr int TRSM_SMEM = (NB * TRSM_ROWS + NB * NB + 2 * NB) * (int)sizeof(double);
__global__ void __launch_bounds__(TRSM_ROWS) k_trsm(PlanDev P, const int32_t* __restrict__ items, int nitems, int jb) {
  const int b = blockIdx.x;
  if (b >= nitems) return;
  const int s = items[2 * b], chunk = items[2 * b + 1];
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int m = (int)(P.rptr[s + 1] - P.rptr[s]);
  const int nb = min(NB, w - jb);
  double* Lp = P.Lx + P.lptr[s];
  B2_DYN_SMEM(raw);
  double* R = reinterpret_cast<double*>(raw);   // [NB][TRSM_ROWS]; its head doubles as the diagonal block S
  double* S = R;                                // [NB x DIAG_LD] column-major (NB*DIAG_LD <= NB*TRSM_ROWS)
  double* Lr = R + NB * TRSM_ROWS;              // [NB][NB] row-major copy of L11 (strict lower)
  double* dsh = Lr + NB * NB;
  double* dd = dsh + NB;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < NB * NB; idx += TRSM_ROWS) {
    const int i = idx % NB, j = idx / NB;
    S[i + j * DIAG_LD] = (i < nb && j <= i) ? Lp[(jb + i) + (size_t)(jb + j) * m] : 0.0;
  }
  __syncthreads();
  cta_ldlt64<TRSM_ROWS>(S, nb, dsh, P.flags);
  if (chunk == 0) {
    double* stage = P.dstage + P.dsptr[s] + (size_t)(jb / NB) * NB * NB;
    for (int idx = tid; idx < NB * NB; idx += TRSM_ROWS) {
      const int i = idx % NB, j = idx / NB;
      if (i < nb && j <= i) stage[i + j * NB] = S[i + j * DIAG_LD];
    }
    if (tid < nb) P.dvec[c0 + jb + tid] = S[tid + tid * DIAG_LD];
  }
  for (int idx = tid; idx < NB * NB; idx += TRSM_ROWS) {
    const int k = idx / NB, t = idx % NB;
    Lr[k * NB + t] = (t < k && k < nb) ? S[k + t * DIAG_LD] : 0.0;
  }
  if (tid < NB) dd[tid] = (tid < nb) ? S[tid + tid * DIAG_LD] : 1.0;
  __syncthreads();                               // S is dead from here on: R takes its place
  const int i = jb + nb + chunk * TRSM_ROWS + tid;
  if (i >= m) return;                            // no barrier below
  for (int k = 0; k < NB; k++) R[k * TRSM_ROWS + tid] = (k < nb) ? Lp[i + (size_t)(jb + k) * m] : 0.0;
  for (int kb = 0; kb < nb; kb += 8) {
    double a8[8];
    B2_UNROLL
    for (int kk = 0; kk < 8; kk++) a8[kk] = R[(kb + kk) * TRSM_ROWS + tid];
    for (int tb = 0; tb < kb; tb += 8) {
      double w8[8];
      B2_UNROLL
      for (int t = 0; t < 8; t++) w8[t] = R[(tb + t) * TRSM_ROWS + tid];
      B2_UNROLL
      for (int kk = 0; kk < 8; kk++) {
        const double* lrow = Lr + (kb + kk) * NB + tb;
        B2_UNROLL
        for (int t = 0; t < 8; t++) a8[kk] -= w8[t] * lrow[t];
      }
    }
    B2_UNROLL
    for (int kk = 1; kk < 8; kk++) {
      const double* lrow = Lr + (kb + kk) * NB + kb;
      B2_UNROLL
      for (int t = 0; t < kk; t++) a8[kk] -= a8[t] * lrow[t];
    }
    B2_UNROLL
    for (int kk = 0; kk < 8; kk++) {
      R[(kb + kk) * TRSM_ROWS + tid] = a8[kk];
      if (kb + kk < nb) Lp[i + (size_t)(jb + kb + kk) * m] = a8[kk] / dd[kb + kk];
    }
  }
}

// item = (front, pivot block): copy a factored diagonal block from the staging area into the
// panel (one launch at the end of the factorization, when nobody reads the unfactored blocks).
__global__ void __launch_bounds__(256) k_diag_writeback(PlanDev P, const int32_t* __restrict__ items, int nitems) {
  const int b = blockIdx.x;
  if (b >= nitems) return;
  const int s = items[2 * b], bi = items[2 * b + 1];
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int m = (int)(P.rptr[s + 1] - P.rptr[s]);
  const int jb = bi * NB, nb = min(NB, w - jb);
  double* Lp = P.Lx + P.lptr[s];
  const double* stage = P.dstage + P.dsptr[s] + (size_t)bi * NB * NB;
  for (int idx = threadIdx.x; idx < NB * NB; idx += 256) {
    const int i = idx % NB, j = idx / NB;
    if (i < nb && j <= i) Lp[(jb + i) + (size_t)(jb + j) * m] = stage[i + j * NB];
  }
}

// item = (front, tile row, tile col), tile row >= tile col.  C -= A diag(d) A'^T on one
// TILE x TILE tile with FP64 tensor-core fragments (8 warps, each a 32 x 16 sub-tile),
// operands = panel columns [k0, k0+K) staged through double-buffered shared memory with a
// register prefetch of the next K-chunk:
//   mode 0 (inside the panel): rows/cols start at org = k0+K, cols < w, C is the panel itself
//          (the next diagonal block is factored by the following k_trsm);
//   mode 1 (contribution block): org = w, cols < m, C is CB (lower triangle), K = w.
__global__ void __launch_bounds__(256, 2) k_update(PlanDev P, const int32_t* __restrict__ items, int nitems, int k0,
                                                int Kreq, int mode) {
  const int b = blockIdx.x;
  if (b >= nitems) return;
  const int s = items[3 * b], ti = items[3 * b + 1], tj = items[3 * b + 2];
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int m = (int)(P.rptr[s + 1] - P.rptr[s]);
  const int r = m - w;
  double* Lp = P.Lx + P.lptr[s];
  int K, org, jend;
  if (mode == 0) { K = min(Kreq, w - k0); org = k0 + K; jend = w; }
  else { k0 = 0; K = w; org = w; jend = m; }
  const int i0 = org + ti * TILE, j0 = org + tj * TILE;
  constexpr int KC = UPD_KC, LDT = TILE + 4;
  __shared__ double As[2][KC][LDT];
  __shared__ double Bs[2][KC][LDT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp & 1, wc = warp >> 1;           // warp tile: rows wr*32.., cols wc*16..
  double acc[4][2][2];
  B2_UNROLL
  for (int a = 0; a < 4; a++)
    B2_UNROLL
    for (int c = 0; c < 2; c++) { acc[a][c][0] = 0.0; acc[a][c][1] = 0.0; }
  // loader: each thread moves KC/4 (k) x 1 (row) elements of A and of B per chunk
  const int lr = tid & 63, lk = tid >> 6;
  const int gi = i0 + lr, gj = j0 + lr;
  const bool vi = gi < m, vj = gj < jend;
  const double* pa = Lp + gi + (size_t)k0 * m;
  const double* pb = Lp + gj + (size_t)k0 * m;
  const double* dv = P.dvec + c0 + k0;
  double ra[KC / 4], rb[KC / 4];
  const int nchunk = (K + KC - 1) / KC;
  auto gload = [&](int kc) {
    B2_UNROLL
    for (int p = 0; p < KC / 4; p++) {
      const int k = kc + lk + 4 * p;
      ra[p] = (vi && k < K) ? pa[(size_t)k * m] : 0.0;
      rb[p] = (vj && k < K) ? pb[(size_t)k * m] * dv[k] : 0.0;
    }
  };
  auto sstore = [&](int buf) {
    B2_UNROLL
    for (int p = 0; p < KC / 4; p++) {
      As[buf][lk + 4 * p][lr] = ra[p];
      Bs[buf][lk + 4 * p][lr] = rb[p];
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  for (int c = 0; c < nchunk; c++) {
    const int buf = c & 1;
    if (c + 1 < nchunk) gload((c + 1) * KC);
    B2_UNROLL
    for (int ks = 0; ks < KC; ks += 4) {
      double af[4], bf[2];
      B2_UNROLL
      for (int a = 0; a < 4; a++) af[a] = As[buf][ks + t][wr * 32 + a * 8 + g];
      B2_UNROLL
      for (int cc = 0; cc < 2; cc++) bf[cc] = Bs[buf][ks + t][wc * 16 + cc * 8 + g];
      B2_UNROLL
      for (int a = 0; a < 4; a++)
        B2_UNROLL
        for (int cc = 0; cc < 2; cc++) dmma_8x8x4(acc[a][cc][0], acc[a][cc][1], af[a], bf[cc]);
    }
    if (c + 1 < nchunk) sstore(buf ^ 1);
    __syncthreads();
  }
  double* Cb = (mode == 0) ? Lp : (P.CB + P.cbptr[s]);
  B2_UNROLL
  for (int a = 0; a < 4; a++)
    B2_UNROLL
    for (int cc = 0; cc < 2; cc++)
      B2_UNROLL
      for (int e = 0; e < 2; e++) {
        const int ri = i0 + wr * 32 + a * 8 + g, cj = j0 + wc * 16 + cc * 8 + 2 * t + e;
        if (ri >= m || cj >= jend || ri < cj) continue;
        double* dst = (mode == 0) ? (Cb + ri + (size_t)cj * m) : (Cb + (ri - w) + (size_t)(cj - w) * r);
        *dst -= acc[a][cc][e];
      }
}

// pivot-sign counts (src/solver_types.jl:90-96): counts[0] = #{d > tol}, [1] = #{|d| <= tol},
// [2] = #{d < -tol}, [3] = #NaN
__global__ void __launch_bounds__(256) k_inertia(const double* __restrict__ d, int64_t n, double tol,
                                                 unsigned long long* __restrict__ counts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int pos = 0, zer = 0, neg = 0, nan = 0;
  for (; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = d[i];
    pos += v > tol;
    zer += fabs(v) <= tol;
    neg += v < -tol;
    nan += (v != v);
  }
  B2_UNROLL
  for (int o = 16; o > 0; o >>= 1) {
    pos += __shfl_xor_sync(0xffffffffu, pos, o);
    zer += __shfl_xor_sync(0xffffffffu, zer, o);
    neg += __shfl_xor_sync(0xffffffffu, neg, o);
    nan += __shfl_xor_sync(0xffffffffu, nan, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (pos) atomicAdd(&counts[0], (unsigned long long)pos);
    if (zer) atomicAdd(&counts[1], (unsigned long long)zer);
    if (neg) atomicAdd(&counts[2], (unsigned long long)neg);
    if (nan) atomicAdd(&counts[3], (unsigned long long)nan);
  }
}

// ------------------------------------------------------------------------------------------
// (3) triangular solves, one CTA per front and one launch per tree level.
// forward:  t = [x(pivots); 0] + sum_children u_c ;  y = L11^{-1} t1 ; u = t2 - L21 y ;
//           x(pivots) <- y / D
// backward: x(pivots) <- L11^{-T} (z - L21^T x(rows below))
// ------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT) k_fwd(PlanDev P, const int32_t* __restrict__ list, int count,
                                            double* __restrict__ x, double* __restrict__ upd) {
  const int b = blockIdx.x;
  if (b >= count) return;
  const int s = list[b];
  B2_DYN_SMEM(raw);
  double* xs = reinterpret_cast<double*>(raw);
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int m = (int)(P.rptr[s + 1] - P.rptr[s]);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* Lp = P.Lx + P.lptr[s];
  for (int i = tid; i < m; i += NT) xs[i] = (i < w) ? x[c0 + i] : 0.0;
  __syncthreads();
  for (int ci = P.child_ptr[s]; ci < P.child_ptr[s + 1]; ci++) {
    const int c = P.child_idx[ci];
    const int wc = P.scol[c + 1] - P.scol[c];
    const int64_t rc0 = P.rptr[c] + wc;
    const int rc = (int)(P.rptr[c + 1] - rc0);
    const int32_t* relc = P.rel + rc0;
    const double* uc = upd + P.uptr[c];
    for (int k = tid; k < rc; k += NT) xs[relc[k]] += uc[k];
    __syncthreads();
  }
  for (int jb = 0; jb < w; jb += SNB) {
    const int nb = min(SNB, w - jb);
    if (warp == 0) {
      double lrow[SNB];
      B2_UNROLL
      for (int k = 0; k < SNB; k++)
        lrow[k] = (lane < nb && k < lane) ? Lp[(jb + lane) + (size_t)(jb + k) * m] : 0.0;
      double y = (lane < nb) ? xs[jb + lane] : 0.0;
      B2_UNROLL
      for (int k = 0; k < SNB; k++) {
        const double yk = __shfl_sync(0xffffffffu, y, k);
        if (lane > k) y -= lrow[k] * yk;
      }
      if (lane < nb) xs[jb + lane] = y;
    }
    __syncthreads();
    for (int i = jb + nb + tid; i < m; i += NT) {
      double acc = 0.0;
      B2_UNROLL
      for (int k = 0; k < SNB; k++)
        if (k < nb) acc += Lp[i + (size_t)(jb + k) * m] * xs[jb + k];
      xs[i] -= acc;
    }
    __syncthreads();
  }
  for (int i = tid; i < w; i += NT) x[c0 + i] = xs[i] / P.dvec[c0 + i];
  double* us = upd + P.uptr[s];
  for (int i = w + tid; i < m; i += NT) us[i - w] = xs[i];
}

template <int NT>
__global__ void __launch_bounds__(NT) k_bwd(PlanDev P, const int32_t* __restrict__ list, int count,
                                            double* __restrict__ x) {
  const int b = blockIdx.x;
  if (b >= count) return;
  const int s = list[b];
  B2_DYN_SMEM(raw);
  double* xs = reinterpret_cast<double*>(raw);
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int64_t r0 = P.rptr[s];
  const int m = (int)(P.rptr[s + 1] - r0);
  double* red = xs + m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  const double* Lp = P.Lx + P.lptr[s];
  for (int i = tid; i < m; i += NT) xs[i] = (i < w) ? x[c0 + i] : x[P.rowidx[r0 + i]];
  __syncthreads();
  const int nblk = (w + SNB - 1) / SNB;
  for (int bi = nblk - 1; bi >= 0; bi--) {
    const int jb = bi * SNB;
    const int nb = min(SNB, w - jb);
    // red[k] = sum_{i >= jb+nb} L(i, jb+k) xs[i]; one warp per column
    for (int k = warp; k < nb; k += NW) {
      const double* col = Lp + (size_t)(jb + k) * m;
      double acc = 0.0;
      for (int i = jb + nb + lane; i < m; i += 32) acc += col[i] * xs[i];
      B2_UNROLL
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) red[k] = acc;
    }
    __syncthreads();
    if (warp == 0) {
      double lcol[SNB];  // lcol[k] = L(jb+k, jb+lane), k > lane
      B2_UNROLL
      for (int k = 0; k < SNB; k++)
        lcol[k] = (lane < nb && k > lane && k < nb) ? Lp[(jb + k) + (size_t)(jb + lane) * m] : 0.0;
      double v = (lane < nb) ? xs[jb + lane] - red[lane] : 0.0;
      B2_UNROLL
      for (int kk = 0; kk < SNB; kk++) {
        const int k = SNB - 1 - kk;
        const double xk = __shfl_sync(0xffffffffu, v, k);
        if (lane < k) v -= lcol[k] * xk;
      }
      if (lane < nb) xs[jb + lane] = v;
    }
    __syncthreads();
  }
  for (int i = tid; i < w; i += NT) x[c0 + i] = xs[i];
}

__global__ void __launch_bounds__(256) k_perm_in(int64_t n, const int32_t* __restrict__ perm,
                                                 const double* __restrict__ b, double* __restrict__ x) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) x[k] = b[perm[k]];
}

// out[perm[k]] = sign * x[k]  (+ optional accumulate for refinement: out += x)
__global__ void __launch_bounds__(256) k_perm_out(int64_t n, const int32_t* __restrict__ perm,
                                                  const double* __restrict__ x, double* __restrict__ out,
                                                  int accumulate) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) {
    const int32_t p = perm[k];
    out[p] = accumulate ? out[p] + x[k] : x[k];
  }
}

__global__ void __launch_bounds__(256) k_scale_copy(int64_t n, const double* __restrict__ in, double* __restrict__ out,
                                                    double alpha) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = alpha * in[k];
}

// res = b - K x with the symmetric CSR (values gathered through the CSC slots); one row/thread
__global__ void __launch_bounds__(256) k_residual(int64_t n, const int64_t* __restrict__ Sp,
                                                  const int32_t* __restrict__ Sj,
                                                  const int32_t* __restrict__ Sslot,
                                                  const double* __restrict__ nzval,
                                                  const double* __restrict__ x, const double* __restrict__ b,
                                                  double* __restrict__ res) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
  for (int64_t q = Sp[i]; q < Sp[i + 1]; q++) acc += nzval[Sslot[q]] * x[Sj[q]];
  res[i] = b[i] - acc;
}

// out[blockIdx.x] = sum of squares of a slice; second launch with one block folds the partials
__global__ void __launch_bounds__(256) k_sumsq(int64_t n, const double* __restrict__ v, double* __restrict__ out) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += v[i] * v[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(256) k_fold(int n, const double* __restrict__ part, double* __restrict__ out) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += part[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

}  // namespace b2
