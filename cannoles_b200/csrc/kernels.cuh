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

// item = (front, destination column block, destination row chunk).  The CTA owns the
// ASM_ROWS x ASM_COLS destination tile in shared memory: zero, scatter the A entries of its
// columns, add the children's contribution blocks one child after the other (fixed order =>
// deterministic sums, no atomics), then write the tile once (panel columns j < w go to Lx, the
// others to CB; only rows >= column are produced).
__global__ void __launch_bounds__(256) k_assemble_large(PlanDev P, const int32_t* __restrict__ items, int nitems) {
  const int b = blockIdx.x;
  if (b >= nitems) return;
  const int s = items[3 * b], j0 = items[3 * b + 1] * ASM_COLS, i0 = items[3 * b + 2] * ASM_ROWS;
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
  for (int ci = P.child_ptr[s]; ci < P.child_ptr[s + 1]; ci++) {
    const int c = P.child_idx[ci];
    const int wc = P.scol[c + 1] - P.scol[c];
    const int64_t rc0 = P.rptr[c] + wc;
    const int rc = (int)(P.rptr[c + 1] - rc0);
    const int32_t* relc = P.rel + rc0;
    const double* cb = P.CB + P.cbptr[c];
    // child columns / rows whose destination falls in the tile: relc is increasing
    int lo = 0, hi = rc;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (relc[mid] < j0) lo = mid + 1; else hi = mid; }
    const int ja = lo;
    hi = rc;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (relc[mid] < je) lo = mid + 1; else hi = mid; }
    const int jz = lo;
    if (ja < jz) {
      lo = ja; hi = rc;   // rows >= column, so the row range starts no earlier than ja
      while (lo < hi) { int mid = (lo + hi) >> 1; if (relc[mid] < i0) lo = mid + 1; else hi = mid; }
      const int ia = lo;
      hi = rc;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (relc[mid] < ie) lo = mid + 1; else hi = mid; }
      const int iz = lo;
      for (int j = ja + warp; j < jz; j += 8) {
        double* dst = &T[relc[j] - j0][0] - i0;
        const double* src = cb + (size_t)j * rc;
        for (int i = max(ia, j) + lane; i < iz; i += 32) dst[relc[i]] += src[i];
      }
    }
    __syncthreads();
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

// Pivot-free LDL^T of an nb x nb (nb <= 32) block stored column-major in shared memory
// (S[i + j * ld], lower triangle), by ONE warp: lane i keeps row i in registers, the unscaled
// column k travels through a small shared buffer (broadcast reads).  On return the strict lower
// part holds L and the diagonal D.  Returns 1 if an exactly zero pivot was met.
__device__ __forceinline__ int warp_ldlt32_smem(double* S, int ld, int nb, double* colbuf, int lane) {
  double a[32];
  B2_UNROLL
  for (int j = 0; j < 32; j++) a[j] = (lane < nb && j <= lane) ? S[lane + j * ld] : 0.0;
  int bad = 0;
  B2_UNROLL
  for (int k = 0; k < 32; k++) {
    if (k < nb) {
      const double dk = __shfl_sync(0xffffffffu, a[k], k);
      const double aik = a[k];
      const double lik = aik / dk;
      if (dk == 0.0) bad = 1;
      double* cbuf = colbuf + (k & 1) * 32;
      cbuf[lane] = aik;
      __syncwarp();
      B2_UNROLL
      for (int j = k + 1; j < 32; j++)
        if (j < nb && lane >= j) a[j] -= lik * cbuf[j];
      if (lane > k) a[k] = lik;
    }
  }
  __syncwarp();
  B2_UNROLL
  for (int j = 0; j < 32; j++)
    if (lane < nb && j <= lane) S[lane + j * ld] = a[j];
  return bad;
}

// CTA-level LDL^T of an nb x nb block, nb <= 64, in shared memory S (ld = DIAG_LD):
// [A11; A21 A22] -> warp LDL^T of A11, triangular solve for A21, Schur update of A22, warp LDL^T of
// A22.  scratch: DIAG_SCRATCH doubles of shared memory.  Every thread of the CTA (NT threads,
// NT >= 32) must call it.
template <int NT>
__device__ __forceinline__ void cta_ldlt64(double* S, int nb, double* scratch, int* flags) {
  constexpr int ld = DIAG_LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n1 = min(nb, 32), n2 = nb - n1;
  if (warp == 0) {
    const int bad = warp_ldlt32_smem(S, ld, n1, scratch, lane);
    if (bad && lane == 0) flags[0] = 1;
  }
  __syncthreads();
  if (n2 > 0) {
    // W = A21 L11^{-T} (kept in registers), L21 = W D1^{-1}; one row per lane of warp 0
    if (warp == 0) {
      double a[32];
      B2_UNROLL
      for (int k = 0; k < 32; k++) a[k] = (lane < n2 && k < n1) ? S[(32 + lane) + k * ld] : 0.0;
      B2_UNROLL
      for (int k = 1; k < 32; k++) {
        double acc = a[k];
        B2_UNROLL
        for (int t = 0; t < k; t++) acc -= a[t] * S[k + t * ld];   // L11(k,t): broadcast read
        a[k] = acc;
      }
      // scratch2[row][k] = W (for the Schur update), S <- L21
      B2_UNROLL
      for (int k = 0; k < 32; k++)
        if (lane < n2 && k < n1) {
          scratch[64 + lane * 33 + k] = a[k];
          S[(32 + lane) + k * ld] = a[k] / S[k + k * ld];
        }
    }
    __syncthreads();
    // A22(i,j) -= sum_k W(i,k) L21(j,k), i >= j
    for (int e = tid; e < 32 * 32; e += NT) {
      const int i = e & 31, j = e >> 5;
      if (i < n2 && j <= i) {
        double acc = 0.0;
        for (int k = 0; k < n1; k++) acc += scratch[64 + i * 33 + k] * S[(32 + j) + k * ld];
        S[(32 + i) + (32 + j) * ld] -= acc;
      }
    }
    __syncthreads();
    if (warp == 0) {
      const int bad = warp_ldlt32_smem(S + 32 + 32 * ld, ld, n2, scratch, lane);
      if (bad && lane == 0) flags[0] = 1;
    }
    __syncthreads();
  }
}

constexpr int DIAG_SCRATCH = 64 + 32 * 33;

// item = (front, row chunk).  Every CTA first factors the nb x nb diagonal block (jb, jb) of the
// panel in shared memory (redundantly: 4 us of work instead of one more launch on the critical
// path; chunk 0 writes the factored block and its pivots back), then forms
// L21 = A21 L11^{-T} D^{-1} for its TRSM_ROWS rows below the block: one row per thread, in two
// halves of 32 columns to bound registers.
__global__ void __launch_bounds__(TRSM_ROWS) k_trsm(PlanDev P, const int32_t* __restrict__ items, int nitems, int jb) {
  const int b = blockIdx.x;
  if (b >= nitems) return;
  const int s = items[2 * b], chunk = items[2 * b + 1];
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int m = (int)(P.rptr[s + 1] - P.rptr[s]);
  const int nb = min(NB, w - jb);
  double* Lp = P.Lx + P.lptr[s];
  __shared__ double L11[NB * DIAG_LD];   // column-major: strict lower = L11, diagonal = D
  __shared__ double scratch[DIAG_SCRATCH];
  __shared__ double dd[NB];
  const int tid = threadIdx.x;
  for (int idx = tid; idx < NB * NB; idx += TRSM_ROWS) {
    const int i = idx % NB, j = idx / NB;
    L11[i + j * DIAG_LD] = (i < nb && j <= i) ? Lp[(jb + i) + (size_t)(jb + j) * m] : 0.0;
  }
  __syncthreads();
  cta_ldlt64<TRSM_ROWS>(L11, nb, scratch, P.flags);
  __syncthreads();
  if (tid < NB) dd[tid] = (tid < nb) ? L11[tid + tid * DIAG_LD] : 1.0;
  if (chunk == 0) {
    for (int idx = tid; idx < NB * NB; idx += TRSM_ROWS) {
      const int i = idx % NB, j = idx / NB;
      if (i < nb && j <= i) Lp[(jb + i) + (size_t)(jb + j) * m] = L11[i + j * DIAG_LD];
    }
    if (tid < nb) P.dvec[c0 + jb + tid] = L11[tid + tid * DIAG_LD];
  }
  __syncthreads();
  const int i = jb + nb + chunk * TRSM_ROWS + tid;
  if (i >= m) return;
  double a[32], a2[32];
  B2_UNROLL
  for (int k = 0; k < 32; k++) a[k] = (k < nb) ? Lp[i + (size_t)(jb + k) * m] : 0.0;
  B2_UNROLL
  for (int k = 1; k < 32; k++) {
    double acc = a[k];
    B2_UNROLL
    for (int t = 0; t < k; t++) acc -= a[t] * L11[k + t * DIAG_LD];
    a[k] = acc;
  }
  if (nb > 32) {
    B2_UNROLL
    for (int k = 0; k < 32; k++) a2[k] = (32 + k < nb) ? Lp[i + (size_t)(jb + 32 + k) * m] : 0.0;
    B2_UNROLL
    for (int k = 0; k < 32; k++) {
      double acc = a2[k];
      B2_UNROLL
      for (int t = 0; t < 32; t++) acc -= a[t] * L11[(32 + k) + t * DIAG_LD];
      B2_UNROLL
      for (int t = 0; t < k; t++) acc -= a2[t] * L11[(32 + k) + (32 + t) * DIAG_LD];
      a2[k] = acc;
    }
    B2_UNROLL
    for (int k = 0; k < 32; k++)
      if (32 + k < nb) Lp[i + (size_t)(jb + 32 + k) * m] = a2[k] / dd[32 + k];
  }
  B2_UNROLL
  for (int k = 0; k < 32; k++)
    if (k < nb) Lp[i + (size_t)(jb + k) * m] = a[k] / dd[k];
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
