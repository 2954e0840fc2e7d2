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

// barrier of the threads that share one front: the whole CTA, or one warp when FPB fronts share a CTA
#define B2_FSYNC() do { if (FPB > 1) __syncwarp(); else __syncthreads(); } while (0)


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

// Validation of a speculative rho retry (Engine::factorize_retry): flag[0] = 1 unless `fresh` (the
// caller's values, just uploaded) equals `prev` (the values of the last upload) bit for bit outside
// the trailing rho segment and holds exactly `rho` inside it.
__global__ void __launch_bounds__(256) k_vals_compare(int64_t n, int64_t nhead, const double* __restrict__ prev,
                                                      const double* __restrict__ fresh, double rho,
                                                      int* __restrict__ flag) {
  const long long rbits = __double_as_longlong(rho);
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const long long f = __double_as_longlong(fresh[i]);
    bad = bad || (i < nhead ? f != __double_as_longlong(prev[i]) : f != rbits);
  }
  if (bad) flag[0] = 1;
}

// ------------------------------------------------------------------------------------------
// (2a) small fronts: the whole m x m front lives in shared memory.
// zero -> scatter A -> extend-add children (fixed order, deterministic) -> eliminate w pivots
// (right-looking, no pivoting) -> write panel, pivots and contribution block.
// ------------------------------------------------------------------------------------------
template <int NT, int FPB>
__global__ void __launch_bounds__(NT * FPB) k_front_small(PlanDev P, const int32_t* __restrict__ list, int count,
                                                          int fstride) {
  // FPB > 1 (the 32-thread class): FPB fronts per CTA, one warp each, warp-level barriers only
  const int slot = (FPB > 1) ? (int)(threadIdx.x >> 5) : 0;
  const int b = blockIdx.x * FPB + slot;
  if (b >= count) return;
  const int s = list[b];
  B2_DYN_SMEM(raw0);
  unsigned char* raw = raw0 + (size_t)slot * fstride;
  double* F = reinterpret_cast<double*>(raw);
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int64_t r0 = P.rptr[s];
  const int m = (int)(P.rptr[s + 1] - r0);
  const int r = m - w;
  double* lk = F + (size_t)m * m;
  double* ak = lk + m;
  const int tid = (FPB > 1) ? (int)(threadIdx.x & 31) : (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;

  for (int idx = tid; idx < m * m; idx += NT) F[idx] = 0.0;
  B2_FSYNC();
  if (P.sf_ptr != nullptr) {
    // extend-add from a flat list built and sorted by destination on the host (A entry first, then
    // the children ascending: the sums of the loop below, bit for bit): the first element of a run
    // sums its run.  Two dependent loads for the whole front instead of a barrier and a chain of
    // five dependent loads per child -- a front of the bottom levels has a dozen one-entry children.
    const int64_t e1 = P.sf_ptr[s + 1];
    for (int64_t e = P.sf_ptr[s] + tid; e < e1; e += NT) {
      const int w0 = P.sf_ent[2 * e];
      const int rl = (int)((unsigned)w0 >> 16);
      if (rl) {
        double sum = 0.0;
        for (int k = 0; k < rl; k++) {
          const int wk = P.sf_ent[2 * (e + k)], src = P.sf_ent[2 * (e + k) + 1];
          sum += (wk & (1 << 15)) ? P.nzval[src] : P.CB[src];
        }
        F[w0 & 0x7fff] = sum;
      }
    }
    B2_FSYNC();
  } else {
  for (int64_t q = P.amap_ptr[s] + tid; q < P.amap_ptr[s + 1]; q += NT)
    F[P.amap_pos[q]] = P.nzval[P.amap_slot[q]];
  B2_FSYNC();
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
    B2_FSYNC();
  }
  }
  // eliminate the w pivot columns: rank-1 updates restricted to the pivot columns ...
  for (int k = 0; k < w; k++) {
    const double dk = F[k + k * m];
    const double rdk = rcp_nr(dk);
    for (int i = k + 1 + tid; i < m; i += NT) {
      const double a = F[i + k * m];
      const double l = a * rdk;
      ak[i] = a;
      lk[i] = l;
      F[i + k * m] = l;
    }
    if (tid == 0) {
      P.dvec[c0 + k] = dk;
      if (dk == 0.0) P.flags[0] = 1;
    }
    B2_FSYNC();
    for (int j = k + 1 + warp; j < w; j += NW) {
      const double ajk = ak[j];
      for (int i = j + lane; i < m; i += 32) F[i + j * m] -= lk[i] * ajk;
    }
    B2_FSYNC();
  }
  // ... then ONE rank-w update of the contribution block, C -= L21 D L21^T, in 8 x 8 x 4 FP64
  // tensor-core tiles straight from shared memory (d_k sits on the diagonal of the panel)
  if (r > 0) {
    const int nt8 = (r + 7) >> 3;
    const int g = lane >> 2, t = lane & 3;
    for (int tile = warp; tile < nt8 * nt8; tile += NW) {
      const int tj = tile / nt8, ti = tile - tj * nt8;
      if (ti < tj) continue;
      const int ra = w + ti * 8 + g, rb = w + tj * 8 + g;
      double acc0 = 0.0, acc1 = 0.0;
      for (int k0 = 0; k0 < w; k0 += 4) {
        const int k = k0 + t;
        double av = 0.0, bv = 0.0;
        if (k < w) {
          if (ra < m) av = F[ra + k * m];
          if (rb < m) bv = F[rb + k * m] * F[k + k * m];
        }
        dmma_8x8x4(acc0, acc1, av, bv);
      }
      const int cj = w + tj * 8 + 2 * t;
      if (ra < m) {
        if (cj < m && ra >= cj) F[ra + cj * m] -= acc0;
        if (cj + 1 < m && ra >= cj + 1) F[ra + (cj + 1) * m] -= acc1;
      }
    }
    B2_FSYNC();
  }
  double* Lp = P.Lx + P.lptr[s];
  for (int idx = tid; idx < m * w; idx += NT) Lp[idx] = F[idx];
  double* cbp = P.CB + P.cbptr[s];
  for (int j = warp; j < r; j += NW)
    for (int i = j + lane; i < r; i += 32) cbp[i + (size_t)j * r] = F[(w + i) + (w + j) * m];
}

// ------------------------------------------------------------------------------------------
// (2a') tiny fronts (order <= MM, MM = 4 or 8): one THREAD per front.  Half of the fronts of a
// KKT system are leaves of order 2-8 (an r_i or lambda_j with its few neighbours): a warp per
// front leaves 31 lanes idle and is CTA-launch bound.  The front lives in a per-thread slice of
// shared memory (element e of thread t at Fs[e * tiny_nt(MM) + t]: conflict-free), same steps as
// k_front_small, no barriers.
// ------------------------------------------------------------------------------------------
template <int MM>
__global__ void __launch_bounds__(tiny_nt(MM)) k_front_tiny(PlanDev P, const int32_t* __restrict__ list, int count) {
  constexpr int TNT = tiny_nt(MM);
  __shared__ double Fs[MM * MM * TNT];
  const int b = blockIdx.x * TNT + threadIdx.x;
  if (b >= count) return;
  const int s = list[b];
  double* F = Fs + threadIdx.x;
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int64_t r0 = P.rptr[s];
  const int m = (int)(P.rptr[s + 1] - r0);
  const int r = m - w;
  for (int e = 0; e < m * m; e++) F[e * TNT] = 0.0;
  for (int64_t q = P.amap_ptr[s]; q < P.amap_ptr[s + 1]; q++) F[P.amap_pos[q] * TNT] = P.nzval[P.amap_slot[q]];
  for (int ci = P.child_ptr[s]; ci < P.child_ptr[s + 1]; ci++) {
    const int c = P.child_idx[ci];
    const int wc = P.scol[c + 1] - P.scol[c];
    const int64_t rc0 = P.rptr[c] + wc;
    const int rc = (int)(P.rptr[c + 1] - rc0);
    const int32_t* relc = P.rel + rc0;
    const double* cb = P.CB + P.cbptr[c];
    for (int j = 0; j < rc; j++) {
      const int J = relc[j];
      for (int i = j; i < rc; i++) F[(relc[i] + J * m) * TNT] += cb[i + (size_t)j * rc];
    }
  }
  for (int k = 0; k < w; k++) {
    const double dk = F[(k + k * m) * TNT];
    P.dvec[c0 + k] = dk;
    if (dk == 0.0) P.flags[0] = 1;
    const double rdk = rcp_nr(dk);
    for (int j = k + 1; j < m; j++) {
      const double lj = F[(j + k * m) * TNT] * rdk;
      for (int i = j; i < m; i++) F[(i + j * m) * TNT] -= F[(i + k * m) * TNT] * lj;   // column k still unscaled
    }
    for (int i = k + 1; i < m; i++) F[(i + k * m) * TNT] *= rdk;
  }
  double* Lp = P.Lx + P.lptr[s];
  for (int j = 0; j < w; j++)
    for (int i = j; i < m; i++) Lp[i + j * m] = F[(i + j * m) * TNT];
  double* cbp = P.CB + P.cbptr[s];
  for (int j = 0; j < r; j++)
    for (int i = j; i < r; i++) cbp[i + (size_t)j * r] = F[((w + i) + (w + j) * m) * TNT];
}

// ------------------------------------------------------------------------------------------
// (2b) tiled path for fronts that do not fit in shared memory.  Pivot blocks of NB = TILE = 64
// columns; per block: diagonal 64 x 64 LDL^T (inside the CTA that produced it) -> k_trsm
// (rows below) -> k_update (trailing pivot columns, FP64 tensor-core tiles); one more k_update
// pass with K = w forms the contribution block.
// ------------------------------------------------------------------------------------------

// item = (front, first destination column, first destination row, global id of that column,
// first / end A entry of the tile's columns relative to amap_ptr[front], tile columns, tile rows).
// The CTA owns an nrows x ncols destination tile in shared memory (16 x 128 for fronts of order
// <= 128, 8 x 256 above): zero, scatter the A
// entries of its columns, add the children's contribution blocks one child after the other
// (fixed order => deterministic sums, no atomics), then write the tile once (panel columns j < w
// go to Lx, the others to CB; only rows >= column are produced).  Which (child, child column)
// pairs land on a destination column and where their data lives is precomputed on the host
// (asm_cptr / asm_ent / asm_rc / asm_off): a front at the top of the tree has hundreds of small
// children and scanning them all per tile was the whole cost of the first version of this kernel.
// (the body is shared with the extend-add tasks of k_front_dag; `it` = the 8 ints of the item, the
// row origin in the low 16 bits of it[2], the tile shape packed in it[6] = ncols | nrows << 16)
__device__ __forceinline__ void assemble_tile_body(const PlanDev& P, const int32_t* __restrict__ it, double* T) {
  const int s = it[0], j0 = it[1], i0 = it[2] & 0xffff;
  const int gcb = it[3];
  const int ncols = it[6] & 0xffff, nrows = it[6] >> 16;   // tile shape: ncols * nrows <= ASM_TILE
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int64_t r0 = P.rptr[s];
  const int m = (int)(P.rptr[s + 1] - r0);
  const int r = m - w;
  const int je = min(j0 + ncols, m), ie = min(i0 + nrows, m);
  // T[(j - j0) * nrows + (i - i0)]: ASM_TILE doubles of shared memory
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int idx = tid; idx < ncols * nrows; idx += 256) T[idx] = 0.0;
  __syncthreads();
  {  // A entries of the tile's pivot columns: range precomputed on the host (amap is sorted by position)
    const int64_t a0 = P.amap_ptr[s];
    const int64_t qa = a0 + it[4], qb = a0 + it[5];
    for (int64_t q = qa + tid; q < qb; q += 256) {
      const int pos = P.amap_pos[q];
      const int j = pos / m, i = pos - j * m;
      if (i >= i0 && i < ie) T[(j - j0) * nrows + (i - i0)] = P.nzval[P.amap_slot[q]];
    }
  }
  __syncthreads();
  // Every warp owns the destination columns j0 + warp, j0 + warp + 8 (< je) of the tile.  For each
  // of them the host has listed, in child order (deterministic sums), the (child, child column)
  // pairs that map onto it: the lanes fetch 32 list entries and the children's descriptors at
  // once, then the warp walks them with shuffles -- no scanning, no dependent pointer chasing,
  // and no barrier between children because no two warps touch the same column of T.
  const bool one_chunk = m <= nrows;
  for (int J = j0 + warp; J < je; J += 8) {
    const int64_t q0 = P.asm_cptr[gcb + (J - j0)], q1 = P.asm_cptr[gcb + (J - j0) + 1];
    double* dst = T + (J - j0) * nrows - i0;
    for (int64_t qb = q0; qb < q1; qb += 32) {
      const int cnt = (int)min((int64_t)32, q1 - qb);
      int jj = 0, rcv = 0;
      long long ro = 0, co = 0;
      if (lane < cnt) {
        const int ce = P.asm_ent[2 * (qb + lane)];
        jj = P.asm_ent[2 * (qb + lane) + 1];
        rcv = P.asm_rc[ce];
        ro = P.asm_off[2 * ce];
        co = P.asm_off[2 * ce + 1];
      }
      for (int k = 0; k < cnt; k++) {
        const int j = __shfl_sync(0xffffffffu, jj, k), rc = __shfl_sync(0xffffffffu, rcv, k);
        const int32_t* relc = P.rel + __shfl_sync(0xffffffffu, ro, k);
        const double* src = P.CB + __shfl_sync(0xffffffffu, co, k) + (size_t)j * rc;
        int ia = j, iz = rc;   // rows >= column
        if (!one_chunk) {
          int lo = j, hi = rc;
          while (lo < hi) { int mid = (lo + hi) >> 1; if (relc[mid] < i0) lo = mid + 1; else hi = mid; }
          ia = lo;
          hi = rc;
          while (lo < hi) { int mid = (lo + hi) >> 1; if (relc[mid] < ie) lo = mid + 1; else hi = mid; }
          iz = lo;
        }
#pragma unroll 4
        for (int i = ia + lane; i < iz; i += 32) dst[relc[i]] += src[i];
        __syncwarp();   // the next (child, column) pair may land on the same rows through other lanes
      }
    }
  }
  __syncthreads();
  double* Lp = P.Lx + P.lptr[s];
  double* cbp = P.CB + P.cbptr[s];
  for (int j = j0 + warp; j < je; j += 8) {
    const double* src = T + (j - j0) * nrows - i0;
    if (j < w) {
      for (int i = max(i0, j) + lane; i < ie; i += 32) Lp[i + (size_t)j * m] = src[i];
    } else {
      for (int i = max(i0, j) + lane; i < ie; i += 32) cbp[(i - w) + (size_t)(j - w) * r] = src[i];
    }
  }
}

__global__ void __launch_bounds__(256) k_assemble_large(PlanDev P, const int32_t* __restrict__ items, int nitems) {
  const int b = blockIdx.x;
  if (b >= nitems) return;
  __shared__ double T[ASM_TILE];
  assemble_tile_body(P, items + 8 * (size_t)b, T);
}

// CTA-level pivot-free LDL^T of an nb x nb block (nb <= NB = 64) held column-major in shared
// memory (S[i + j * DIAG_LD], lower triangle), blocked in 8-column panels.
// Warp 0 owns the pivot chain: EVERY lane factors the 8 x 8 diagonal sub-block redundantly in
// registers (no shuffles, no shared memory on the chain), two pivots per step: with the current
// 2 x 2 leading block [a b; b c] the Schur complement of both pivots is
//     x_ij -= (p_i u_j + q_i v_j) / det,   p = c u - b v,  q = a v - b u,  det = a c - b^2,
// where (u, v) are the two columns below the block; everything but the product with 1 / det is
// formed while the reciprocal is in flight.  D stays diagonal: d_k = a, d_{k+1} = det / a,
// l_{i,k} = u_i / a, l_{i,k+1} = q_i / det -- the same L D L^T as the scalar elimination.
// Everything else is spread over the CTA (a single warp retires ~1 instruction every 4 cycles,
// measured, so whatever stays on warp 0 is the critical path):
//   (B) substitution of the rows below the 8 x 8 block, one row per thread (L to S, W = L D to Wd);
//   (C) rank-8 update of the trailing block by warps 1.., while warp 0 updates only the NEXT
//       8 x 8 diagonal block and goes straight on to factor it (look-ahead inside the CTA).
// Two barriers per panel.  Wd: NB x 8 + 8 doubles of shared memory.
// On return strict lower = L, diagonal = D.
__device__ __forceinline__ double rcp_cubic(double d) {
#ifdef B2_EMULATE
  return 1.0 / d;
#else
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  const double e = __fma_rn(-d, x, 1.0);      // ~2^-20
  const double t = __fma_rn(e, e, e);         // e + e^2
  return __fma_rn(x, t, x);                   // x (1 + e + e^2): relative error e^3
#endif
}

// warp 0: factor the 8 x 8 block at (kb, kb) (pw <= 8 valid columns) in registers, store L8 / D
// back into S and the reciprocal pivots into rds
// (not inlined: one copy per kernel instead of one per call site -- the fully unrolled body is ~9 KB
// of SASS and the code around the pivot chain runs cold, instruction fetch included)
// (measured and rejected in round 2: spreading the 8 x 8 block over the lanes -- two entries per lane,
// nine 64-bit shuffles per 2 x 2 pivot step -- takes 2440 cycles per panel against 1816 for this
// redundant form: scripts/_dev/ldlt_microbench.cu, 28.3 k vs 24.3 k cycles per 64 x 64 tile)
__device__ __noinline__ void warp_ldlt8(double* S, int kb, int pw, double* rds, int* flags) {
  constexpr int ld = DIAG_LD;
  double g[8][8], rd[8];
  B2_UNROLL
  for (int c = 0; c < 8; c++)
    B2_UNROLL
    for (int t = 0; t <= c; t++) g[c][t] = (c < pw) ? S[(kb + c) + (kb + t) * ld] : (c == t ? 1.0 : 0.0);
  bool bad = false;
  B2_UNROLL
  for (int k = 0; k < 8; k += 2) {
    const double a = g[k][k], b = g[k + 1][k], c = g[k + 1][k + 1];
    const double det = __fma_rn(a, c, -(b * b));
    bad = bad || a == 0.0 || det == 0.0;
    const double rdet = rcp_cubic(det);
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
  // every lane holds the same values: same-address stores, one of them lands
  B2_UNROLL
  for (int c = 0; c < 8; c++) {
    if (c < pw) {
      B2_UNROLL
      for (int t = 0; t <= c; t++) S[(kb + c) + (kb + t) * ld] = g[c][t];
      rds[c] = rd[c];
    }
  }
  if (bad && (threadIdx.x & 31) == 0) flags[0] = 1;
}

template <int NT>
__device__ __forceinline__ void cta_ldlt64(double* S, int nb, double* Wd, int* flags) {
  constexpr int ld = DIAG_LD;
  static_assert(NT >= 64, "cta_ldlt64 needs a pivot warp and at least one update warp");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* rds = Wd + NB * 8;
  if (warp == 0) warp_ldlt8(S, 0, min(8, nb), rds, flags);
  for (int kb = 0; kb < nb; kb += 8) {
#ifdef B2_TIMING
    long long tp0 = clock64();
#endif
    const int pw = min(8, nb - kb);
    const int t0 = kb + 8;
    __syncthreads();                    // L8 / 1/d of this panel are in S / rds; the trailing block is up to date
    // (B) rows below the 8 x 8 block: w[c] = a[c] - sum_{t<c} w[t] L8[c][t],  l[c] = w[c] / d_c
    for (int row = t0 + tid; row < nb; row += NT) {
      double wv[8];
      B2_UNROLL
      for (int c = 0; c < 8; c++) wv[c] = (c < pw) ? S[row + (kb + c) * ld] : 0.0;
      B2_UNROLL
      for (int c = 1; c < 8; c++)
        B2_UNROLL
        for (int t = 0; t < c; t++) wv[c] -= wv[t] * ((c < pw) ? S[(kb + c) + (kb + t) * ld] : 0.0);
      B2_UNROLL
      for (int c = 0; c < 8; c++) {
        if (c < pw) { S[row + (kb + c) * ld] = wv[c] * rds[c]; Wd[row * 8 + c] = wv[c]; }
        else Wd[row * 8 + c] = 0.0;
      }
    }
    __syncthreads();                    // panel kb complete
    B2_ACC(10, tp0);
#ifdef B2_TIMING
    tp0 = clock64();
#endif
    if (t0 >= nb) break;
    if (warp == 0) {
      // the next 8 x 8 diagonal block only: lane <-> (row i, columns 2 jp, 2 jp + 1), then its LDL^T
      const int i = t0 + (lane >> 2), j = t0 + 2 * (lane & 3);
      if (i < nb) {
        double s0 = 0.0, s1 = 0.0;
        B2_UNROLL
        for (int c = 0; c < 8; c++) {
          const double l = S[i + (kb + c) * ld];
          s0 += l * Wd[j * 8 + c];
          s1 += l * Wd[(j + 1) * 8 + c];     // j + 1 <= t0 + 7 < NB: in bounds, unused when beyond the row
        }
        if (j <= i) S[i + j * ld] -= s0;
        if (j + 1 <= i) S[i + (j + 1) * ld] -= s1;
      }
      __syncwarp();
      warp_ldlt8(S, t0, min(8, nb - t0), rds, flags);
    } else {
      // (C) rows t0 + 8 .. nb - 1, columns t0 .. row.  Thread <-> (row pair, column phase): rows
      // t1 + p and nb - 1 - p together have the same number of columns whatever p, so the
      // trapezoid is balanced; the phases (one per warp) take every NPH-th column.
      const int t1 = t0 + 8;
      const int n = nb - t1;
      const int p = lane, phase = warp - 1;
      constexpr int NPH = NT / 32 - 1;
      const int iA = t1 + p, iB = nb - 1 - p;
      if (n > 0 && p < (n + 1) / 2) {
        const bool two = iB > iA;
        double la[8], lb[8];
        B2_UNROLL
        for (int c = 0; c < 8; c++) {
          la[c] = S[iA + (kb + c) * ld];
          lb[c] = two ? S[iB + (kb + c) * ld] : 0.0;
        }
        const int jmax = two ? iB : iA;
#pragma unroll 2
        for (int j = t0 + phase; j <= jmax; j += NPH) {
          const double* wj = Wd + j * 8;
          double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
          B2_UNROLL
          for (int c = 0; c < 8; c += 2) {
            const double w0 = wj[c], w1 = wj[c + 1];
            a0 += la[c] * w0; a1 += la[c + 1] * w1;
            b0 += lb[c] * w0; b1 += lb[c + 1] * w1;
          }
          if (j <= iA) S[iA + j * ld] -= a0 + a1;
          if (two) S[iB + j * ld] -= b0 + b1;
        }
      }
    }
    B2_ACC(11, tp0);
  }
}

// Look-ahead factorization of the diagonal block (jb, jb) of the listed fronts, one CTA each:
// unless `first`, the block is first brought up to date with the previous pivot block,
// A -= L(blk, prev) D L(blk, prev)^T (k_update leaves that region alone), then factored
// (cta_ldlt64) into the staging area + dvec.  It runs on a side stream concurrently with the
// trailing update of the previous block, so that the serial 64-pivot chain is off the critical
// path of k_trsm.  Dynamic shared memory: DIAG_SMEM bytes.
constexpr int DIAG_SMEM = (NB * DIAG_LD + 2 * NB * (NB + 1) + 2 * NB * 8) * (int)sizeof(double);
__global__ void __launch_bounds__(256) k_diag(PlanDev P, const int32_t* __restrict__ items, int nitems, int jb, int first) {
  const int b = blockIdx.x;
  if (b >= nitems) return;
  const int s = items[b];
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int m = (int)(P.rptr[s + 1] - P.rptr[s]);
  if (jb >= w) return;
  const int nb = min(NB, w - jb);
  const double* Lp = P.Lx + P.lptr[s];
  B2_DYN_SMEM(raw);
  double* S = reinterpret_cast<double*>(raw);    // [NB x DIAG_LD] column-major
  double* As = S + NB * DIAG_LD;                 // [k][i], ld NB + 1: L(jb + i, k0 + k)
  double* Ws = As + NB * (NB + 1);               // [k][j]: L(jb + j, k0 + k) d(k0 + k)
  double* Wd = Ws + NB * (NB + 1);               // scratch of cta_ldlt64
  const int tid = threadIdx.x;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int i = idx % NB, j = idx / NB;
    S[i + j * DIAG_LD] = (i < nb && j <= i) ? Lp[(jb + i) + (size_t)(jb + j) * m] : 0.0;
  }
  if (!first) {
    const int k0 = jb - NB;                      // the previous pivot block is always full
    for (int idx = tid; idx < NB * NB; idx += 256) {
      const int i = idx % NB, k = idx / NB;
      const double l = (i < nb) ? Lp[(jb + i) + (size_t)(k0 + k) * m] : 0.0;
      As[k * (NB + 1) + i] = l;
      Ws[k * (NB + 1) + i] = l * P.dvec[c0 + k0 + k];
    }
    __syncthreads();
    // thread <-> (row i, column phase): columns j = phase, phase + 4, ... <= i
    const int i = tid & 63, phase = tid >> 6;
    if (i < nb) {
      for (int j = phase; j <= i; j += 4) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll 8
        for (int k = 0; k < NB; k += 2) {
          a0 += As[k * (NB + 1) + i] * Ws[k * (NB + 1) + j];
          a1 += As[(k + 1) * (NB + 1) + i] * Ws[(k + 1) * (NB + 1) + j];
        }
        S[i + j * DIAG_LD] -= a0 + a1;
      }
    }
  }
  __syncthreads();
  cta_ldlt64<256>(S, nb, Wd, P.flags);
  double* stage = P.dstage + P.dsptr[s] + (size_t)(jb / NB) * NB * NB;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int i = idx % NB, j = idx / NB;
    if (i < nb && j <= i) stage[i + j * NB] = S[i + j * DIAG_LD];
  }
  if (tid < nb) P.dvec[c0 + jb + tid] = S[tid + tid * DIAG_LD];
}

// item = (front, row chunk).  Every CTA first factors the nb x nb diagonal block (jb, jb) of the
// panel in shared memory (redundantly: a few us of work instead of one more launch on the
// critical path; chunk 0 stores the factored block in the staging area -- NOT in the panel, which
// sibling CTAs may still be reading -- and the pivots in dvec), then forms
// L21 = A21 L11^{-T} D^{-1} for its TRSM_ROWS rows below the block: TRSM_RPT rows per thread, rows staged in shared memory, substitution in
// 8-column blocks (rolled loops, 8 x 8 unrolled bodies).
// Dynamic shared memory: TRSM_SMEM bytes.
constexpr int TRSM_SMEM = (NB * TRSM_ROWS + NB * NB + 2 * NB * 8 + NB) * (int)sizeof(double);
__global__ void __launch_bounds__(TRSM_THREADS) k_trsm(PlanDev P, const int32_t* __restrict__ items, int nitems, int jb,
                                                       int prefact) {
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
  double* Wd = Lr + NB * NB;                    // [2][NB][8] scratch of cta_ldlt64
  double* dd = Wd + 2 * NB * 8;
  const int tid = threadIdx.x;
#ifdef B2_TIMING
  if (tid == 0 && blockIdx.x == 0) { b2_dbg[10] = 0; b2_dbg[11] = 0; }
#endif
  B2_TICK(0);
  double* stage = P.dstage + P.dsptr[s] + (size_t)(jb / NB) * NB * NB;
  if (prefact) {
    // look-ahead mode: k_diag has already factored this block (concurrently with the previous
    // trailing update); it is in the staging area
    for (int idx = tid; idx < NB * NB; idx += TRSM_THREADS) {
      const int i = idx % NB, j = idx / NB;
      S[i + j * DIAG_LD] = (i < nb && j <= i) ? stage[i + j * NB] : 0.0;
    }
    __syncthreads();
  } else {
    for (int idx = tid; idx < NB * NB; idx += TRSM_THREADS) {
      const int i = idx % NB, j = idx / NB;
      S[i + j * DIAG_LD] = (i < nb && j <= i) ? Lp[(jb + i) + (size_t)(jb + j) * m] : 0.0;
    }
    __syncthreads();
    B2_TICK(1);
    cta_ldlt64<TRSM_THREADS>(S, nb, Wd, P.flags);
    B2_TICK(2);
    if (chunk == 0) {
      for (int idx = tid; idx < NB * NB; idx += TRSM_THREADS) {
        const int i = idx % NB, j = idx / NB;
        if (i < nb && j <= i) stage[i + j * NB] = S[i + j * DIAG_LD];
      }
      if (tid < nb) P.dvec[c0 + jb + tid] = S[tid + tid * DIAG_LD];
    }
  }
  for (int idx = tid; idx < NB * NB; idx += TRSM_THREADS) {
    const int k = idx / NB, t = idx % NB;
    Lr[k * NB + t] = (t < k && k < nb) ? S[k + t * DIAG_LD] : 0.0;
  }
  if (tid < NB) dd[tid] = (tid < nb) ? rcp_nr(S[tid + tid * DIAG_LD]) : 1.0;   // reciprocal pivots
  __syncthreads();                               // S is dead from here on: R takes its place
  B2_TICK(3);
  // TRSM_RPT rows per thread (tid, tid + TRSM_THREADS, ... of the chunk), two sets of partial sums
  // per row: 16 * TRSM_RPT independent FMA chains
  const int ibase = jb + nb + chunk * TRSM_ROWS;
  if (ibase + tid >= m) return;                  // no barrier below
  int irow[TRSM_RPT];
  bool vrow[TRSM_RPT];
  B2_UNROLL
  for (int h = 0; h < TRSM_RPT; h++) { irow[h] = ibase + h * TRSM_THREADS + tid; vrow[h] = irow[h] < m; }
  for (int k = 0; k < NB; k++) {
    B2_UNROLL
    for (int h = 0; h < TRSM_RPT; h++)
      R[k * TRSM_ROWS + h * TRSM_THREADS + tid] = (k < nb && vrow[h]) ? Lp[irow[h] + (size_t)(jb + k) * m] : 0.0;
  }
  B2_TICK(4);
  for (int kb = 0; kb < nb; kb += 8) {
    double a8[TRSM_RPT][8], b8[TRSM_RPT][8];
    B2_UNROLL
    for (int kk = 0; kk < 8; kk++)
      B2_UNROLL
      for (int h = 0; h < TRSM_RPT; h++) {
        a8[h][kk] = R[(kb + kk) * TRSM_ROWS + h * TRSM_THREADS + tid];
        b8[h][kk] = 0.0;
      }
    for (int tb = 0; tb < kb; tb += 8) {
      double w8[TRSM_RPT][8];
      B2_UNROLL
      for (int t = 0; t < 8; t++)
        B2_UNROLL
        for (int h = 0; h < TRSM_RPT; h++) w8[h][t] = R[(tb + t) * TRSM_ROWS + h * TRSM_THREADS + tid];
      B2_UNROLL
      for (int kk = 0; kk < 8; kk++) {
        const double* lrow = Lr + (kb + kk) * NB + tb;
        B2_UNROLL
        for (int t = 0; t < 8; t += 2) {
          const double l0 = lrow[t], l1 = lrow[t + 1];
          B2_UNROLL
          for (int h = 0; h < TRSM_RPT; h++) {
            a8[h][kk] -= w8[h][t] * l0;
            b8[h][kk] -= w8[h][t + 1] * l1;
          }
        }
      }
    }
    B2_UNROLL
    for (int kk = 0; kk < 8; kk++)
      B2_UNROLL
      for (int h = 0; h < TRSM_RPT; h++) a8[h][kk] += b8[h][kk];
    B2_UNROLL
    for (int kk = 1; kk < 8; kk++) {
      const double* lrow = Lr + (kb + kk) * NB + kb;
      B2_UNROLL
      for (int t = 0; t < kk; t++) {
        const double l = lrow[t];
        B2_UNROLL
        for (int h = 0; h < TRSM_RPT; h++) a8[h][kk] -= a8[h][t] * l;
      }
    }
    B2_UNROLL
    for (int kk = 0; kk < 8; kk++) {
      const double rdk = dd[(kb + kk) & (NB - 1)];
      B2_UNROLL
      for (int h = 0; h < TRSM_RPT; h++) {
        R[(kb + kk) * TRSM_ROWS + h * TRSM_THREADS + tid] = a8[h][kk];
        if (kb + kk < nb && vrow[h]) Lp[irow[h] + (size_t)(jb + kb + kk) * m] = a8[h][kk] * rdk;
      }
    }
  }
  B2_TICK(5);
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
// TMA = true: the operand tiles arrive by cp.async.bulk + mbarrier instead of LDG -> registers -> STS
// (selected with B2_UPDATE_TMA=1; measured on C3, 66 k tiles of K = 9 .. 27 ... 249: 1.64 ms against
// 1.50 ms for the register-staged pipeline, which therefore stays the default -- the tiles are 16
// columns of 512 bytes per chunk, too small for the copy engine to beat 256 threads loading them)
template <bool TMA>
__global__ void __launch_bounds__(256, 2) k_update(PlanDev P, const int32_t* __restrict__ items, int nitems, int k0,
                                                   int Kreq, int mode, int skipdiag) {
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
  const int nbnext = (mode == 0) ? min(NB, w - org) : 0;   // rows of the next pivot block: [org, org + nbnext)
  constexpr int KC = UPD_KC, LDT = TILE + 4;
  __shared__ __align__(16) double As[2][KC][LDT];
  __shared__ __align__(16) double Bs[2][KC][LDT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp & 1, wc = warp >> 1;           // warp tile: rows wr*32.., cols wc*16..
  double acc[4][2][2];
  B2_UNROLL
  for (int a = 0; a < 4; a++)
    B2_UNROLL
    for (int c = 0; c < 2; c++) { acc[a][c][0] = 0.0; acc[a][c][1] = 0.0; }
  const int nchunk = (K + KC - 1) / KC;
  const double* dv = P.dvec + c0 + k0;
  double* Cb = (mode == 0) ? Lp : (P.CB + P.cbptr[s]);
  double cold[4][2][2];
  auto load_c = [&]() {   // the C tile is read early (its latency hides behind the K loop) and written once at the end
    B2_UNROLL
    for (int a = 0; a < 4; a++)
      B2_UNROLL
      for (int cc = 0; cc < 2; cc++)
        B2_UNROLL
        for (int e = 0; e < 2; e++) {
          const int ri = i0 + wr * 32 + a * 8 + g, cj = j0 + wc * 16 + cc * 8 + 2 * t + e;
          const bool ok = ri < m && cj < jend && ri >= cj;
          const double* src = (mode == 0) ? (Cb + ri + (size_t)cj * m) : (Cb + (ri - w) + (size_t)(cj - w) * r);
          cold[a][cc][e] = ok ? *src : 0.0;
        }
  };
#ifndef B2_EMULATE
  if constexpr (TMA) {
  // ---- operand pipeline by TMA: one cp.async.bulk per column of the A and of the B tile (64 rows = 512
  // bytes, contiguous in the column-major panel) straight into shared memory, completion on an mbarrier
  // per buffer; nothing is staged in registers.  A column starts at Lp + i0 + k m, which is 16-byte
  // aligned only when that offset is even: the copy starts at the even address at or below it and is
  // 66 doubles long, and the fragment loads add the column's shift (0 or 1).  The rows a partial tile
  // does not have (and the one or two doubles around the 64) are whatever sits there in the panel: they
  // only reach rows / columns of C that are not stored.  Columns k >= K of the last chunk are masked
  // in the fragment loads.  d_k goes to shared memory by plain loads and scales the B fragments.
  __shared__ __align__(8) unsigned long long bar[2];
  __shared__ double ds[2][KC];
  const long long offA = (long long)P.lptr[s] + i0 + (long long)k0 * m, offB = (long long)P.lptr[s] + j0 + (long long)k0 * m;
  const int mpar = m & 1;
  const int shA0 = (int)(offA & 1), shB0 = (int)(offB & 1);
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  __syncthreads();
  auto issue = [&](int c) {   // warp 0: lanes 0..15 the A columns of chunk c, lanes 16..31 the B columns
    const int buf = c & 1, kc = c * KC;
    const int nk = min(KC, K - kc);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the generic-proxy reads of this buffer are behind us
    if (lane == 0) mbar_expect_tx(&bar[buf], (uint32_t)(2 * nk * 66 * sizeof(double)));
    __syncwarp();
    const int kk = lane & 15;
    if (kk < nk) {
      const bool isB = lane >= 16;
      const long long off = (isB ? offB : offA) + (long long)(kc + kk) * m;
      const double* src = P.Lx + (off & ~1LL);
      double* dst = isB ? &Bs[buf][kk][0] : &As[buf][kk][0];
      tma_bulk_g2s(dst, src, (uint32_t)(66 * sizeof(double)), &bar[buf]);
    }
  };
  auto load_d = [&](int c) {
    if (tid < KC) { const int k = c * KC + tid; ds[c & 1][tid] = k < K ? dv[k] : 0.0; }
  };
  B2_TICK(20);
  if (warp == 0) issue(0);
  load_d(0);
  load_c();
  __syncthreads();                       // ds[0]
  B2_TICK(21);
  uint32_t ph0 = 0, ph1 = 0;
  for (int c = 0; c < nchunk; c++) {
    const int buf = c & 1, kc = c * KC;
    if (c + 1 < nchunk) {                // buffer buf ^ 1 was released by the barrier that ended chunk c - 1
      if (warp == 0) issue(c + 1);
      load_d(c + 1);
    }
    mbar_wait(&bar[buf], buf ? ph1 : ph0);
    if (buf) ph1 ^= 1; else ph0 ^= 1;
    B2_UNROLL
    for (int ks = 0; ks < KC; ks += 4) {
      const int k = ks + t;
      const bool kv = kc + k < K;
      const int sa = (shA0 + (kc + k) * mpar) & 1, sb = (shB0 + (kc + k) * mpar) & 1;
      const double dk = ds[buf][k];
      double af[4], bf[2];
      B2_UNROLL
      for (int a = 0; a < 4; a++) af[a] = kv ? As[buf][k][wr * 32 + a * 8 + g + sa] : 0.0;
      B2_UNROLL
      for (int cc = 0; cc < 2; cc++) bf[cc] = kv ? Bs[buf][k][wc * 16 + cc * 8 + g + sb] * dk : 0.0;
      B2_UNROLL
      for (int a = 0; a < 4; a++)
        B2_UNROLL
        for (int cc = 0; cc < 2; cc++) dmma_8x8x4(acc[a][cc][0], acc[a][cc][1], af[a], bf[cc]);
    }
    __syncthreads();                     // chunk c has been read by everyone: its buffer may be refilled; ds[buf ^ 1] is complete
  }
  } else
#endif
  {
  // loader: each thread moves KC/4 (k) x 1 (row) elements of A and of B per chunk
  const int lr = tid & 63, lk = tid >> 6;
  const int gi = i0 + lr, gj = j0 + lr;
  const bool vi = gi < m, vj = gj < jend;
  const double* pa = Lp + gi + (size_t)k0 * m;
  const double* pb = Lp + gj + (size_t)k0 * m;
  double ra[KC / 4], rb[KC / 4], rk[KC / 4];
  // the loads only: the product with d_k is formed when the registers go to shared memory, one
  // chunk of tensor work later (a multiply here would wait for the loads it is meant to hide)
  auto gload = [&](int kc) {
    B2_UNROLL
    for (int p = 0; p < KC / 4; p++) {
      const int k = kc + lk + 4 * p;
      ra[p] = (vi && k < K) ? pa[(size_t)k * m] : 0.0;
      rb[p] = (vj && k < K) ? pb[(size_t)k * m] : 0.0;
      rk[p] = (k < K) ? dv[k] : 0.0;
    }
  };
  auto sstore = [&](int buf) {
    B2_UNROLL
    for (int p = 0; p < KC / 4; p++) {
      As[buf][lk + 4 * p][lr] = ra[p];
      Bs[buf][lk + 4 * p][lr] = rb[p] * rk[p];
    }
  };
  B2_TICK(20);
  gload(0);
  load_c();
  sstore(0);
  __syncthreads();
  B2_TICK(21);
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
  }
  B2_TICK(22);
  B2_UNROLL
  for (int a = 0; a < 4; a++)
    B2_UNROLL
    for (int cc = 0; cc < 2; cc++)
      B2_UNROLL
      for (int e = 0; e < 2; e++) {
        const int ri = i0 + wr * 32 + a * 8 + g, cj = j0 + wc * 16 + cc * 8 + 2 * t + e;
        if (ri >= m || cj >= jend || ri < cj) continue;
        // look-ahead mode: the next diagonal block is brought up to date and factored by k_diag
        if (skipdiag && ri < org + nbnext) continue;
        double* dst = (mode == 0) ? (Cb + ri + (size_t)cj * m) : (Cb + (ri - w) + (size_t)(cj - w) * r);
        *dst = cold[a][cc][e] - acc[a][cc][e];
      }
  B2_TICK(23);
}

// ------------------------------------------------------------------------------------------
// (2d) Dataflow factorization of the tiled fronts of one tree level: ONE launch replaces the
// k_trsm / k_update / k_diag_writeback launch chain (two launches per 64-column pivot block on
// the critical path).  The front is cut into NB x NB tiles: row/column blocks 0 .. np-1 are the
// pivot blocks, np .. np+nq-1 the blocks of the contribution block (which starts at row w,
// not at a multiple of NB).  A plain task is one tile (I, J), I >= J, left-looking:
//     C(I,J) -= sum_{p < min(J, np)} L(I,p) D_p L(J,p)^T        (FP64 DMMA, operands via L2)
//     J < np, I == J : pivot-free LDL^T of the tile (cta_ldlt64), write L/D in place, publish
//                      the inverses of its eight 8 x 8 unit-lower diagonal blocks + 1/d
//     J < np, I >  J : L(I,J) = C L(J,J)^{-T} D^{-1} by block substitution in DMMA fragments
//     J >= np        : the tile of the contribution block is complete
// The tiles (J, J-1) and (J, J) of a pivot block J >= 1 form ONE task (a "chain" task): both are
// brought up to date with the pivot blocks p < J-1 ahead of time; when block J-1 is factored the
// task substitutes (J, J-1), applies that last update to (J, J) straight from shared memory and
// factors it -- the chain diag(J-1) -> diag(J) stays on one SM, with one flag hop.
// Every tile of a pivot column has a flag in global memory, raised when its final L is
// visible; a task waits per p for flag(I,p) and flag(J,p), so the pivot block p+1 is factored
// as soon as ITS tiles are up to date while the rest of the trailing update is still running
// on other SMs (look-ahead without a scheduler).  CTAs take tasks from a ticket counter in a
// host-built topological order (column by column): a CTA only waits for tasks with a smaller
// ticket, which are running or done, so the waits cannot deadlock whatever the number of
// resident CTAs.  The accumulation order per tile is fixed (p ascending): results do not depend
// on the schedule.  The diagonal tile of a chain task is brought up to date with p < J-1 by a
// third kind of task ("ypre", on another SM) that hands it over through the panel.
// task item = (front, I, J | kind << 16, first flag of the front), kind 0 plain, 1 chain, 2 ypre.
// ------------------------------------------------------------------------------------------
#ifdef B2_TIMING
constexpr int DAG_TRACE_MAX = 1 << 16;
constexpr int DAG_TRACE_W = 12;
__device__ long long b2_dag_trace[DAG_TRACE_W * DAG_TRACE_MAX];
__device__ __forceinline__ long long dag_now() { long long v; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v)); return v; }
#define DAG_TRACE(slot) do { if (tid == 0 && trace_base + tk < DAG_TRACE_MAX) b2_dag_trace[DAG_TRACE_W * (trace_base + tk) + (slot)] = dag_now(); } while (0)
#else
#define DAG_TRACE(slot)
#endif
constexpr int DAG_LDT = TILE + 4;
constexpr int DAG_LDL = NB + 1;
constexpr int DAG_SMEM = (4 * UPD_KC * DAG_LDT + NB * DAG_LDT + NB * DAG_LDL + 8 * 64 + NB) * (int)sizeof(double);
constexpr int DAG_SMEM_EXCL = 120 * 1024;   // more than half an SM's shared memory: one CTA per SM

__device__ __forceinline__ int dag_peek(const int* f) {
#ifdef B2_EMULATE
  return *f;
#else
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
  return v;
#endif
}
__device__ __forceinline__ int dag_peek_relaxed(const int* f) {
#ifdef B2_EMULATE
  return *f;
#else
  return *reinterpret_cast<const volatile int*>(f);
#endif
}
__device__ __forceinline__ void dag_wait(const int* f) {
#ifdef B2_EMULATE
  // the emulator runs the tasks one after the other in ticket order: the producer must be done
  if (*f == 0) { fprintf(stderr, "k_front_dag: wait on a task that has not run\n"); abort(); }
#else
  // poll with plain volatile loads (an acquire load invalidates the SM's L1 every time: CCTL.IVALL on
  // the load/store pipe the neighbour's shared-memory traffic goes through), acquire once at the end
  while (dag_peek_relaxed(f) == 0) __nanosleep(64);
  (void)dag_peek(f);
#endif
}
__device__ __forceinline__ void dag_raise(int* f) {
#ifdef B2_EMULATE
  *f = 1;
#else
  __threadfence();
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(f), "r"(1) : "memory");
#endif
}
// counters (extend-add tasks done per 64-column group of a front, contribution-block tiles done per
// front, children done per front): a waiter needs every increment's writes, hence fence - add - fence
__device__ __forceinline__ void dag_wait_count(const int* c, int expect) {
#ifdef B2_EMULATE
  if (*c < expect) { fprintf(stderr, "k_front_dag: wait on a counter that is not complete (%d < %d)\n", *c, expect); abort(); }
#else
  while (dag_peek_relaxed(c) < expect) __nanosleep(64);
  (void)dag_peek(c);
#endif
}
__device__ __forceinline__ int dag_count_up(int* c) {
#ifdef B2_EMULATE
  return (*c)++;
#else
  __threadfence();
  const int old = atomicAdd(c, 1);
  __threadfence();
  return old;
#endif
}
// operands written by other CTAs of the same launch: read through L2
__device__ __forceinline__ double dag_ld(const double* p) {
#ifdef B2_EMULATE
  return *p;
#else
  return __ldcg(p);
#endif
}
// C fragment (c0, c1 = C[g][2t], C[g][2t+1]) -> A fragment of its K-step h: A[g][4h + t]
__device__ __forceinline__ double frag_c2a(double c0, double c1, int h, int lane) {
  const int src = (lane & ~3) | (2 * h + ((lane & 3) >> 1));
  const double v0 = __shfl_sync(0xffffffffu, c0, src), v1 = __shfl_sync(0xffffffffu, c1, src);
  return (lane & 1) ? v1 : v0;
}

// Extend-add of ONE NB x NB tile of a front, straight into the registers that hold the tile as C
// fragments (cold[a][cc][e] <-> row wr*32 + a*8 + g, column wc*16 + cc*8 + 2t + e).
// A separator front has one tiny child per pivot column (the r / lambda leaves hanging off its
// variables: ~80 children land on a diagonal tile, one or two entries each) and two big ones.
//  * big children are listed per tile (tl_ptr / tl_ent, children ascending).  An entry is
//    self-contained -- order of the child's contribution block, its offset, and for the 64 rows and
//    the 64 columns of the tile the child row / column that lands there (uint16, 0xffff: none) -- so
//    every thread gathers ITS 16 entries by destination: no shared-memory tile, no barrier between
//    children, all loads of all children independent of each other;
//  * the A entries and everything that comes from small children (order of the contribution block
//    <= DAG_SMALL_RC) are a FLAT list per tile, built and sorted by destination on the host
//    (fl_ptr / fl_ent: destination | A flag | run length, source offset): one parallel gather into a
//    staging area, then the first element of every run sums its run in list order (A entry first,
//    children ascending) into a shared-memory tile -- two barriers whatever the number of children;
//    a tile without such entries (most tiles away from the diagonal) skips this part.
// Sum per entry: big children ascending, then + (A + small children ascending).  Fixed order, no
// atomics: deterministic.
constexpr int DAG_ENT = 72;                      // ints per tl_ent entry: 8 header + 64 (128 uint16 maps)
constexpr int DAG_DESC = 3;                      // entries staged per round
constexpr int DAG_STAGE = 4096;                  // elements of the flat list gathered per round
constexpr int DAG_SMALL_RC = 32;
__device__ __forceinline__ void dag_assemble_tile(const PlanDev& P, int tgid, double* Ts, double* stage, int* sdesc,
                                                  double (&cold)[4][2][2]) {
  constexpr int LDT = DAG_LDT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3, wr = warp & 1, wc = warp >> 1;
  const int e0 = P.tl_ptr[tgid], e1 = P.tl_ptr[tgid + 1];
  const int f0 = P.fl_ptr[tgid], f1 = P.fl_ptr[tgid + 1];
  B2_UNROLL
  for (int a = 0; a < 4; a++)
    B2_UNROLL
    for (int cc = 0; cc < 2; cc++) { cold[a][cc][0] = 0.0; cold[a][cc][1] = 0.0; }
  if (f1 > f0)
    for (int idx = tid; idx < NB * LDT; idx += 256) Ts[idx] = 0.0;
  for (int eb = e0; eb < e1; eb += DAG_DESC) {
    const int cnt = min(DAG_DESC, e1 - eb);
    // the descriptors of the previous round -- or of the previous CALL: a chain task of pivot block 1
    // assembles tile (1, 0) and then tile (1, 1) with nothing in between -- have been read by everyone
    // (found in round 2: without this barrier a fast warp overwrote them under a slow one, one
    // factorization in ~100 of the C3 slice came out different; tests/test_gpu_parity.py loops on it now)
    __syncthreads();
    if (tid < DAG_ENT * cnt) sdesc[tid] = P.tl_ent[DAG_ENT * (size_t)eb + tid];      // DAG_ENT * DAG_DESC <= 256
    __syncthreads();
    for (int k = 0; k < cnt; k++) {
      const int* d = sdesc + DAG_ENT * k;
      const int rc = d[0];
      const double* cb = P.CB + (((long long)d[2] << 32) | (unsigned)d[1]);
      const unsigned short* rmap = reinterpret_cast<const unsigned short*>(d + 8);
      const unsigned short* cmap = rmap + 64;
      int ri[4], cj[2][2];
      B2_UNROLL
      for (int a = 0; a < 4; a++) ri[a] = rmap[wr * 32 + a * 8 + g];
      B2_UNROLL
      for (int cc = 0; cc < 2; cc++) { cj[cc][0] = cmap[wc * 16 + cc * 8 + 2 * t]; cj[cc][1] = cmap[wc * 16 + cc * 8 + 2 * t + 1]; }
      double v[4][2][2];
      B2_UNROLL
      for (int a = 0; a < 4; a++)
        B2_UNROLL
        for (int cc = 0; cc < 2; cc++)
          B2_UNROLL
          for (int e = 0; e < 2; e++) {
            const bool ok = ri[a] != 0xffff && cj[cc][e] != 0xffff && ri[a] >= cj[cc][e];
            v[a][cc][e] = ok ? cb[ri[a] + (size_t)cj[cc][e] * rc] : 0.0;
          }
      B2_UNROLL
      for (int a = 0; a < 4; a++)
        B2_UNROLL
        for (int cc = 0; cc < 2; cc++) { cold[a][cc][0] += v[a][cc][0]; cold[a][cc][1] += v[a][cc][1]; }
    }
  }
  if (f1 <= f0) return;
  for (int fb = f0; fb < f1; fb += DAG_STAGE) {
    const int n = min(DAG_STAGE, f1 - fb);
    const int32_t* fe = P.fl_ent + 2 * (size_t)fb;
    for (int e = tid; e < n; e += 256) {
      const int w0 = fe[2 * e], w1 = fe[2 * e + 1];
      stage[e] = (w0 & (1 << 13)) ? P.nzval[w1] : P.CB[w1];
    }
    __syncthreads();                             // (also: the tile is zeroed / the previous round is summed)
    for (int e = tid; e < n; e += 256) {
      const int w0 = fe[2 * e];
      const int rl = w0 >> 14;
      if (rl) {
        double sum = Ts[w0 & 0x1fff];
        for (int k = 0; k < rl; k++) sum += stage[e + k];
        Ts[w0 & 0x1fff] = sum;
      }
    }
    __syncthreads();
  }
  B2_UNROLL
  for (int a = 0; a < 4; a++)
    B2_UNROLL
    for (int cc = 0; cc < 2; cc++)
      B2_UNROLL
      for (int e = 0; e < 2; e++) cold[a][cc][e] += Ts[(wc * 16 + cc * 8 + 2 * t + e) * LDT + wr * 32 + a * 8 + g];
  __syncthreads();                               // Ts (= As | Bs) goes back to the operand pipeline
}

__global__ void __launch_bounds__(256, 2) k_front_dag(PlanDev P, const int32_t* __restrict__ items, int ntasks,
                                                      int* __restrict__ tflag, int* __restrict__ ticket, int trace_base) {
  constexpr int KC = UPD_KC, LDT = DAG_LDT, LDL = DAG_LDL;
  B2_DYN_SMEM(raw);
  double* As = reinterpret_cast<double*>(raw);   // [2][KC][LDT]
  double* Bs = As + 2 * KC * LDT;                // [2][KC][LDT]   (As | Bs = NB x LDT: W of a chain task)
  double* T = Bs + 2 * KC * LDT;                 // the tile after the update: [col][row] ld LDT (substitution) / [row + col * DIAG_LD] (LDL^T)
  double* Lr = T + NB * LDT;                     // L(J,J) row-major, ld LDL (substitution) / scratch of cta_ldlt64
  double* Mi = Lr + NB * LDL;                    // inverses of the 8 x 8 diagonal blocks of L(J,J), row-major
  double* rdv = Mi + 8 * 64;                     // 1 / d of the pivot block
  __shared__ int s_tk;
  __shared__ int sdesc[DAG_ENT * DAG_DESC];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp & 1, wc = warp >> 1;       // warp tile of the update: rows wr*32.., cols wc*16..
  const int lr = tid & 63, lk = tid >> 6;
  for (;;) {
    __syncthreads();                             // the previous task is done with shared memory and s_tk
    if (tid == 0) s_tk = atomicAdd(ticket, 1);
    __syncthreads();
    const int tk = s_tk;
    if (tk >= ntasks) break;
    const int32_t* it = items + 8 * (size_t)tk;
    const int s = it[0], I = it[1], fb = it[3];
    const int J = it[2] & 0xffff;
    const int kind = it[2] >> 16;
    const bool chain = kind == 1;                // I == J >= 1: tiles (J, J-1) and (J, J)
    const bool ypre = kind == 2;                 // I == J >= 2: tile (J, J) up to date with the pivot blocks p < J-1
    // task = (front, I, J | kind << 16, first flag, first tile list, first A range, -, front record);
    // per-front record: [0] -, [1] tiles of its contribution block, [2] parent's record (-1: none in
    // this launch), [3] children in this launch
    const int df = it[7];
    const int32_t* fr = P.dfr + 4 * (size_t)df;
#ifdef B2_TIMING
    if (tid == 0 && trace_base + tk < DAG_TRACE_MAX) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      b2_dag_trace[DAG_TRACE_W * (trace_base + tk) + 0] = ((long long)s << 32) | (chain ? (1 << 30) : 0) | (I << 15) | J;
      b2_dag_trace[DAG_TRACE_W * (trace_base + tk) + 1] = smid;
    }
#endif
    DAG_TRACE(2);
    const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
    const int m = (int)(P.rptr[s + 1] - P.rptr[s]);
    const int r = m - w;
    const int np = (w + NB - 1) / NB, nrb = np + (r + NB - 1) / NB;
    const int i0 = I < np ? I * NB : w + (I - np) * NB, iend = I < np ? min(i0 + NB, w) : min(i0 + NB, m);
    const bool pivcol = J < np;
    double* Lp = P.Lx + P.lptr[s];
    double acc[4][2][2];
    // the children factored by THIS launch (the others are complete: earlier launches) hand their
    // contribution blocks over through a counter of the front
    if (fr[3] > 0) {
      if (tid == 0) dag_wait_count(P.dcnt + P.dcnt_ch + df, fr[3]);
      __syncthreads();
    }
    DAG_TRACE(10);
    // ---- the update(s): one tile, or (J, J-1) then (J, J) of a chain task ------------------
    for (int ph = 0; ph < (chain ? 2 : 1); ph++) {
      const int Jt = (chain && ph == 0) ? J - 1 : J;            // column block of this tile
      const int j0 = Jt < np ? Jt * NB : w + (Jt - np) * NB, jend = Jt < np ? min(j0 + NB, w) : min(j0 + NB, m);
      // pivot columns applied here (a chain task takes its diagonal tile from the ypre task)
      const int Ktot = chain ? (ph == 0 ? (J - 1) * NB : 0) : (ypre ? (J - 1) * NB : (pivcol ? J * NB : w));
      const bool from_ypre = chain && ph == 1 && J >= 2;
      if (from_ypre) {
        if (tid == 0) dag_wait(tflag + fb + np * nrb + J);
        __syncthreads();
      }
      const int* fI = tflag + fb + I;            // flag of tile (I, p): fI[p * nrb]
      const int* fJ = tflag + fb + Jt;
      B2_UNROLL
      for (int a = 0; a < 4; a++)
        B2_UNROLL
        for (int c = 0; c < 2; c++) { acc[a][c][0] = 0.0; acc[a][c][1] = 0.0; }
      // the tile before any update: the extend-add happens HERE, in shared memory (As | Bs, free at
      // this point) -- the A entries of the tile, then the children's contribution blocks one child
      // after the other (fixed order, no atomics: deterministic sums) -- so an assembled front is never
      // written to / read back from HBM; a chain task takes its diagonal tile from the ypre task
      double cold[4][2][2];
      if (from_ypre) {
        B2_UNROLL
        for (int a = 0; a < 4; a++)
          B2_UNROLL
          for (int cc = 0; cc < 2; cc++)
            B2_UNROLL
            for (int e = 0; e < 2; e++) {
              const int ri = i0 + wr * 32 + a * 8 + g, cj = j0 + wc * 16 + cc * 8 + 2 * t + e;
              const bool ok = ri < iend && cj < jend && ri >= cj;
              cold[a][cc][e] = ok ? dag_ld(Lp + ri + (size_t)cj * m) : 0.0;
            }
      } else {
        dag_assemble_tile(P, it[4] + Jt * nrb + I, As, Lr, sdesc, cold);
        if (ph == 0) DAG_TRACE(11);
      }
      const int nchunk = (Ktot + KC - 1) / KC;
      if (nchunk > 0) {
        const int gi = i0 + lr, gj = j0 + lr;
        const bool vi = gi < iend, vj = gj < jend;
        const double* pa = Lp + gi;
        const double* pb = Lp + gj;
        const double* dv = P.dvec + c0;
        double ra[KC / 4], rb[KC / 4], rk[KC / 4];
        // the loads only: the product with d_k is formed when the registers go to shared memory
        auto gload = [&](int kc) {
          B2_UNROLL
          for (int p = 0; p < KC / 4; p++) {
            const int k = kc + lk + 4 * p;
            ra[p] = (vi && k < Ktot) ? dag_ld(pa + (size_t)k * m) : 0.0;
            rb[p] = (vj && k < Ktot) ? dag_ld(pb + (size_t)k * m) : 0.0;
            rk[p] = (k < Ktot) ? dag_ld(dv + k) : 0.0;
          }
        };
        auto sstore = [&](int buf) {
          B2_UNROLL
          for (int p = 0; p < KC / 4; p++) {
            As[(buf * KC + lk + 4 * p) * LDT + lr] = ra[p];
            Bs[(buf * KC + lk + 4 * p) * LDT + lr] = rb[p] * rk[p];
          }
        };
        if (tid == 0) { dag_wait(fI); dag_wait(fJ); }
        __syncthreads();
        gload(0);
        sstore(0);
        __syncthreads();
        for (int c = 0; c < nchunk; c++) {
          const int buf = c & 1;
          const bool more = c + 1 < nchunk;
          const int kn = (c + 1) * KC;
          bool pre = more;
          if (more && (kn % NB) == 0) {          // the next chunk starts a new pivot block: is it there yet?
            const int pn = kn / NB;
            pre = __syncthreads_and(dag_peek_relaxed(fI + pn * nrb) != 0 && dag_peek_relaxed(fJ + pn * nrb) != 0) != 0;
            if (pre) { (void)dag_peek(fI + pn * nrb); (void)dag_peek(fJ + pn * nrb); }
          }
          if (pre) gload(kn);
          B2_UNROLL
          for (int ks = 0; ks < KC; ks += 4) {
            double af[4], bf[2];
            B2_UNROLL
            for (int a = 0; a < 4; a++) af[a] = As[(buf * KC + ks + t) * LDT + wr * 32 + a * 8 + g];
            B2_UNROLL
            for (int cc = 0; cc < 2; cc++) bf[cc] = Bs[(buf * KC + ks + t) * LDT + wc * 16 + cc * 8 + g];
            B2_UNROLL
            for (int a = 0; a < 4; a++)
              B2_UNROLL
              for (int cc = 0; cc < 2; cc++) dmma_8x8x4(acc[a][cc][0], acc[a][cc][1], af[a], bf[cc]);
          }
          if (more) {
            if (!pre) {                          // not yet: wait now that the chunk in hand is consumed
              const int pn = kn / NB;
              if (tid == 0) { dag_wait(fI + pn * nrb); dag_wait(fJ + pn * nrb); }
              __syncthreads();
              gload(kn);
            }
            sstore(buf ^ 1);
          }
          __syncthreads();
        }
      }
      // from here on acc = -(updated tile): the original tile is folded in (frees its registers)
      B2_UNROLL
      for (int a = 0; a < 4; a++)
        B2_UNROLL
        for (int cc = 0; cc < 2; cc++) { acc[a][cc][0] -= cold[a][cc][0]; acc[a][cc][1] -= cold[a][cc][1]; }
      if (pivcol && I > Jt) {                    // rows below a pivot block: to shared memory for the substitution
        B2_UNROLL
        for (int a = 0; a < 4; a++)
          B2_UNROLL
          for (int cc = 0; cc < 2; cc++)
            B2_UNROLL
            for (int e = 0; e < 2; e++) {
              const int li = wr * 32 + a * 8 + g, lj = wc * 16 + cc * 8 + 2 * t + e;
              T[lj * LDT + li] = -acc[a][cc][e];
            }
      }
    }
    DAG_TRACE(3);
    if (ypre) {                                  // hand the partially updated diagonal tile over through the panel
      const int j0 = J * NB, jend = min(j0 + NB, w);
      B2_UNROLL
      for (int a = 0; a < 4; a++)
        B2_UNROLL
        for (int cc = 0; cc < 2; cc++)
          B2_UNROLL
          for (int e = 0; e < 2; e++) {
            const int ri = i0 + wr * 32 + a * 8 + g, cj = j0 + wc * 16 + cc * 8 + 2 * t + e;
            if (ri < iend && cj < jend && ri >= cj) Lp[ri + (size_t)cj * m] = -acc[a][cc][e];
          }
      __syncthreads();
      if (tid == 0) dag_raise(tflag + fb + np * nrb + J);
      DAG_TRACE(9);
      continue;
    }
    if (!pivcol) {                               // contribution block: done
      const int j0 = w + (J - np) * NB, jend = min(j0 + NB, m);
      double* Cb = P.CB + P.cbptr[s];
      B2_UNROLL
      for (int a = 0; a < 4; a++)
        B2_UNROLL
        for (int cc = 0; cc < 2; cc++)
          B2_UNROLL
          for (int e = 0; e < 2; e++) {
            const int ri = i0 + wr * 32 + a * 8 + g, cj = j0 + wc * 16 + cc * 8 + 2 * t + e;
            if (ri < iend && cj < jend && ri >= cj) Cb[(ri - w) + (size_t)(cj - w) * r] = -acc[a][cc][e];
          }
      // the last tile of the contribution block completes the front: tell the parent
      __syncthreads();
      if (tid == 0 && dag_count_up(P.dcnt + P.dcnt_cb + df) == fr[1] - 1 && fr[2] >= 0)
        (void)dag_count_up(P.dcnt + P.dcnt_ch + fr[2]);
      DAG_TRACE(9);
      continue;
    }
    if (chain || I > J) {
      // ---- substitution against pivot block Js: W = C L(Js,Js)^{-T} in 8-column blocks, right-
      // looking inside the tile; warp <-> 8 rows, the eight 8 x 8 blocks of its rows live in C
      // fragments.  L(I,Js) = W D^{-1} goes to the panel; a chain task also keeps W (in As|Bs) and
      // W D^{-1} (in T) for the update of its diagonal tile.
      const int Js = chain ? J - 1 : J;
      const int js0 = Js * NB, nbs = min(NB, w - js0);
      const double* stage = P.dstage + P.dsptr[s] + (size_t)Js * NB * NB;
      int* fS = tflag + fb + Js * nrb;           // flags of column Js
      if (tid == 0) dag_wait(fS + Js);
      __syncthreads();
      DAG_TRACE(4);
      for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx % NB, j = idx / NB;
        Lr[i * LDL + j] = (j < i && i < nbs) ? dag_ld(Lp + (js0 + i) + (size_t)(js0 + j) * m) : 0.0;
      }
      for (int idx = tid; idx < 8 * 64; idx += 256) Mi[idx] = dag_ld(stage + idx);
      if (tid < NB) rdv[tid] = dag_ld(stage + 512 + tid);
      __syncthreads();
      DAG_TRACE(5);
      {
        // warp <-> 8 rows; left-looking over the 8-column blocks with ROLLED loops (this code runs
        // once per task, cold: a fully unrolled right-looking version was instruction-fetch bound).
        // W of the finished blocks sits in Ws (= As|Bs, [col][row]); two accumulator pairs halve
        // the dependent DMMA chain.
        double* Ws = As;
        const int row = warp * 8 + g;
        const int gi = i0 + row;
        for (int kb = 0; kb < 8; kb++) {
          double s0 = 0.0, s1 = 0.0, u0 = 0.0, u1 = 0.0;
          const double* lrow = Lr + (kb * 8 + g) * LDL + t;
          for (int tb = 0; tb < kb; tb++) {
            dmma_8x8x4(s0, s1, Ws[(tb * 8 + t) * LDT + row], lrow[tb * 8]);
            dmma_8x8x4(u0, u1, Ws[(tb * 8 + 4 + t) * LDT + row], lrow[tb * 8 + 4]);
          }
          const double c0v = T[(kb * 8 + 2 * t) * LDT + row] - (s0 + u0);
          const double c1v = T[(kb * 8 + 2 * t + 1) * LDT + row] - (s1 + u1);
          const double a0 = frag_c2a(c0v, c1v, 0, lane), a1 = frag_c2a(c0v, c1v, 1, lane);
          double w0 = 0.0, w1 = 0.0;
          dmma_8x8x4(w0, w1, a0, Mi[kb * 64 + g * 8 + t]);
          dmma_8x8x4(w0, w1, a1, Mi[kb * 64 + g * 8 + 4 + t]);
          B2_UNROLL
          for (int e = 0; e < 2; e++) {
            const int lc = kb * 8 + 2 * t + e;
            const double wv = e ? w1 : w0, lv = wv * rdv[lc];
            if (gi < iend && lc < nbs) Lp[gi + (size_t)(js0 + lc) * m] = lv;
            Ws[lc * LDT + row] = wv;
            if (chain) T[lc * LDT + row] = lv;
          }
          __syncwarp();                          // W of block kb is read back as A fragments by the other lanes
        }
      }
      __syncthreads();
      DAG_TRACE(6);
      if (!chain) {
        if (tid == 0) dag_raise(fS + I);
        DAG_TRACE(9);
        continue;
      }
      // warps 4 and 6 own the sub-tiles strictly above the diagonal of (J, J): nothing to update
      if (tid == 6 * 32) dag_raise(fS + I);
      if (warp != 4 && warp != 6) {
        for (int ks = 0; ks < NB; ks += 4) {
          double af[4], bf[2];
          B2_UNROLL
          for (int a = 0; a < 4; a++) af[a] = As[(ks + t) * LDT + wr * 32 + a * 8 + g];
          B2_UNROLL
          for (int cc = 0; cc < 2; cc++) bf[cc] = T[(ks + t) * LDT + wc * 16 + cc * 8 + g];
          B2_UNROLL
          for (int a = 0; a < 4; a++)
            B2_UNROLL
            for (int cc = 0; cc < 2; cc++) dmma_8x8x4(acc[a][cc][0], acc[a][cc][1], af[a], bf[cc]);
        }
      }
      __syncthreads();                           // everybody is done reading T
    }
    // ---- pivot-free LDL^T of the diagonal tile (J, J) ----------------------------------------
    {
      const int j0 = J * NB, nb = min(NB, w - j0);
      double* stage = P.dstage + P.dsptr[s] + (size_t)J * NB * NB;
      B2_UNROLL
      for (int a = 0; a < 4; a++)
        B2_UNROLL
        for (int cc = 0; cc < 2; cc++)
          B2_UNROLL
          for (int e = 0; e < 2; e++) {
            const int li = wr * 32 + a * 8 + g, lj = wc * 16 + cc * 8 + 2 * t + e;
            if (li >= lj) T[li + lj * DIAG_LD] = -acc[a][cc][e];
          }
      __syncthreads();
      DAG_TRACE(7);
      cta_ldlt64<256>(T, nb, Lr, P.flags);
      DAG_TRACE(8);
      for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx % NB, j = idx / NB;
        if (i < nb && j <= i) Lp[(j0 + i) + (size_t)(j0 + j) * m] = T[i + j * DIAG_LD];
      }
      if (tid < nb) P.dvec[c0 + j0 + tid] = T[tid + tid * DIAG_LD];
      if (tid < 64) {
        // column jc of the inverse of the unit-lower 8 x 8 diagonal block kb (identity beyond nb)
        const int kb = tid >> 3, jc = tid & 7;
        double x[8];
        B2_UNROLL
        for (int i = 0; i < 8; i++) x[i] = (i == jc) ? 1.0 : 0.0;
        B2_UNROLL
        for (int i = 1; i < 8; i++) {
          double sum = 0.0;
          B2_UNROLL
          for (int tt = 0; tt < i; tt++) {
            const double l = (kb * 8 + i < nb) ? T[(kb * 8 + i) + (kb * 8 + tt) * DIAG_LD] : 0.0;
            sum += l * x[tt];
          }
          if (i > jc) x[i] = -sum;
        }
        B2_UNROLL
        for (int i = 0; i < 8; i++) stage[kb * 64 + i * 8 + jc] = x[i];
      } else if (tid < 128) {
        const int c = tid - 64;
        stage[512 + c] = (c < nb) ? rcp_nr(T[c + c * DIAG_LD]) : 1.0;
      }
      __syncthreads();
      if (tid == 0) dag_raise(tflag + fb + J * nrb + J);
      DAG_TRACE(9);
    }
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
template <int NT, int FPB>
__global__ void __launch_bounds__(NT * FPB) k_fwd(PlanDev P, const int32_t* __restrict__ list, int count,
                                                  double* __restrict__ x, double* __restrict__ upd, int fstride) {
  // FPB > 1 (the 32-thread class): FPB fronts per CTA, one warp each, warp-level barriers only
  const int slot = (FPB > 1) ? (int)(threadIdx.x >> 5) : 0;
  const int b = blockIdx.x * FPB + slot;
  if (b >= count) return;
  const int s = list[b];
  B2_DYN_SMEM(raw0);
  unsigned char* raw = raw0 + (size_t)slot * fstride;
  double* xs = reinterpret_cast<double*>(raw);
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int m = (int)(P.rptr[s + 1] - P.rptr[s]);
  const int tid = (FPB > 1) ? (int)(threadIdx.x & 31) : (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* Lp = P.Lx + P.lptr[s];
  // own entries of the right-hand side + the children's update vectors: a per-row gather built on
  // the host (ug_ptr / ug_src, children ascending => the sums of the per-child loop, bit for bit),
  // one thread per row, independent loads, no barrier per child
  {
    const int32_t* gp = P.ug_ptr + P.rptr[s];
    for (int i = tid; i < m; i += NT) {
      double acc = (i < w) ? x[c0 + i] : 0.0;
      const int e1 = gp[i + 1];
#pragma unroll 4
      for (int e = gp[i]; e < e1; e++) acc += upd[P.ug_src[e]];
      xs[i] = acc;
    }
  }
  B2_FSYNC();
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
        if (k >= nb) break;   // uniform: most fronts are narrow (w <= 8), 32 shuffle steps would be wasted
        const double yk = __shfl_sync(0xffffffffu, y, k);
        if (lane > k) y -= lrow[k] * yk;
      }
      if (lane < nb) xs[jb + lane] = y;
    }
    B2_FSYNC();
    for (int i = jb + nb + tid; i < m; i += NT) {
      double acc = 0.0;
      B2_UNROLL
      for (int k = 0; k < SNB; k++)
        if (k < nb) acc += Lp[i + (size_t)(jb + k) * m] * xs[jb + k];
      xs[i] -= acc;
    }
    B2_FSYNC();
  }
  for (int i = tid; i < w; i += NT) x[c0 + i] = xs[i] / P.dvec[c0 + i];
  double* us = upd + P.uptr[s];
  for (int i = w + tid; i < m; i += NT) us[i - w] = xs[i];
}

template <int NT, int FPB>
__global__ void __launch_bounds__(NT * FPB) k_bwd(PlanDev P, const int32_t* __restrict__ list, int count,
                                                  double* __restrict__ x, int fstride) {
  // FPB > 1 (the 32-thread class): FPB fronts per CTA, one warp each, warp-level barriers only
  const int slot = (FPB > 1) ? (int)(threadIdx.x >> 5) : 0;
  const int b = blockIdx.x * FPB + slot;
  if (b >= count) return;
  const int s = list[b];
  B2_DYN_SMEM(raw0);
  unsigned char* raw = raw0 + (size_t)slot * fstride;
  double* xs = reinterpret_cast<double*>(raw);
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int64_t r0 = P.rptr[s];
  const int m = (int)(P.rptr[s + 1] - r0);
  double* red = xs + m;
  const int tid = (FPB > 1) ? (int)(threadIdx.x & 31) : (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  const double* Lp = P.Lx + P.lptr[s];
  for (int i = tid; i < m; i += NT) xs[i] = (i < w) ? x[c0 + i] : x[P.rowidx[r0 + i]];
  B2_FSYNC();
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
    B2_FSYNC();
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
    B2_FSYNC();
  }
  for (int i = tid; i < w; i += NT) x[c0 + i] = xs[i];
}

// (3a') tiny fronts: one thread per front (see k_front_tiny)
template <int MM>
__global__ void __launch_bounds__(tiny_nt(MM)) k_fwd_tiny(PlanDev P, const int32_t* __restrict__ list, int count,
                                                      double* __restrict__ x, double* __restrict__ upd) {
  constexpr int TNT = tiny_nt(MM);
  __shared__ double Xs[MM * TNT];
  const int b = blockIdx.x * TNT + threadIdx.x;
  if (b >= count) return;
  const int s = list[b];
  double* xs = Xs + threadIdx.x;
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int m = (int)(P.rptr[s + 1] - P.rptr[s]);
  const double* Lp = P.Lx + P.lptr[s];
  // the children's update entries of a front are contiguous in ug_src (row by row, children
  // ascending within a row): one flat loop whose loads are independent of each other
  for (int i = 0; i < m; i++) xs[i * TNT] = (i < w) ? x[c0 + i] : 0.0;
  {
    const int32_t* gp = P.ug_ptr + P.rptr[s];
    const int e1 = gp[m];
#pragma unroll 4
    for (int e = gp[0]; e < e1; e++) xs[P.ug_row[e] * TNT] += upd[P.ug_src[e]];
  }
  for (int j = 0; j < w; j++) {
    const double yj = xs[j * TNT];
    for (int i = j + 1; i < m; i++) xs[i * TNT] -= Lp[i + j * m] * yj;
    x[c0 + j] = yj / P.dvec[c0 + j];
  }
  double* us = upd + P.uptr[s];
  for (int i = w; i < m; i++) us[i - w] = xs[i * TNT];
}

template <int MM>
__global__ void __launch_bounds__(tiny_nt(MM)) k_bwd_tiny(PlanDev P, const int32_t* __restrict__ list, int count,
                                                      double* __restrict__ x) {
  constexpr int TNT = tiny_nt(MM);
  __shared__ double Xs[MM * TNT];
  const int b = blockIdx.x * TNT + threadIdx.x;
  if (b >= count) return;
  const int s = list[b];
  double* xs = Xs + threadIdx.x;
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int64_t r0 = P.rptr[s];
  const int m = (int)(P.rptr[s + 1] - r0);
  const double* Lp = P.Lx + P.lptr[s];
  for (int i = 0; i < m; i++) xs[i * TNT] = (i < w) ? x[c0 + i] : x[P.rowidx[r0 + i]];
  for (int j = w - 1; j >= 0; j--) {
    double acc = xs[j * TNT];
    for (int i = j + 1; i < m; i++) acc -= Lp[i + j * m] * xs[i * TNT];
    xs[j * TNT] = acc;
    x[c0 + j] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// Explicit inverse X = L^{-1} of a unit-lower 64 x 64 diagonal block of a big front, item =
// (front, block), one CTA of 256 threads, blocked in 8 x 8 sub-blocks:
//   X_ii = L_ii^{-1}                                  (64 threads: one column of one block each)
//   X_ij = -X_ii sum_{j <= k < i} L_ik X_kj,  j < i   block row after block row (7 steps); inside a
//          step every thread forms ONE entry of T = L_i,: X_:,j (<= 56 terms), then of -X_ii T
// Rows beyond the block are identity rows.  Runs once per factorization; the solves read X instead
// of L (a 63-step row-by-row version with one thread per column took 63 us per launch: every step a
// barrier and a dependent chain).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_linv(PlanDev P, const int32_t* __restrict__ items, int nitems) {
  const int b = blockIdx.x;
  if (b >= nitems) return;
  const int s = items[2 * b], c = items[2 * b + 1];
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int m = (int)(P.rptr[s + 1] - P.rptr[s]);
  const int i0 = c * SB, nrow = min(SB, w - i0);
  const double* Lp = P.Lx + P.lptr[s];
  double* out = P.Linv + (size_t)(P.linv_idx[s] + c) * SB * SB;
  constexpr int LD = SB + 1;
  __shared__ double X[SB * LD];      // [row][col]: L on entry (unit lower, identity rows beyond nrow), X = L^{-1} in place
  __shared__ double T[8 * LD];       // [row of the block row][col]
  const int tid = threadIdx.x;
  for (int e = tid; e < SB * SB; e += 256) {
    const int i = e % SB, k = e / SB;          // coalesced over the rows of a column
    double v = (i == k) ? 1.0 : 0.0;
    if (i < nrow && k < i) v = Lp[(i0 + i) + (size_t)(i0 + k) * m];
    X[i * LD + k] = v;
  }
  __syncthreads();
  {                                            // diagonal blocks: column jc of the inverse of block kb
    const int kb = (tid & 63) >> 3, jc = tid & 7, o = kb * 8;
    double x[8], l[8][8];
    B2_UNROLL
    for (int i = 1; i < 8; i++)
      B2_UNROLL
      for (int t = 0; t < i; t++) l[i][t] = X[(o + i) * LD + o + t];
    B2_UNROLL
    for (int i = 0; i < 8; i++) x[i] = (i == jc) ? 1.0 : 0.0;
    B2_UNROLL
    for (int i = 1; i < 8; i++) {
      double sum = 0.0;
      B2_UNROLL
      for (int t = 0; t < i; t++) sum += l[i][t] * x[t];
      if (i > jc) x[i] = -sum;
    }
    __syncthreads();                           // every block has been read before anyone overwrites it
    if (tid < SB) {
      B2_UNROLL
      for (int i = 0; i < 8; i++) X[(o + i) * LD + o + jc] = x[i];
    }
  }
  __syncthreads();
  for (int ib = 1; ib < 8; ib++) {             // block row ib: columns 0 .. 8 ib - 1 (still L there; X above and on the diagonal)
    const int ncol = 8 * ib;
    for (int e = tid; e < 8 * ncol; e += 256) {
      const int r = e / ncol, cc = e - r * ncol;     // a warp: consecutive columns of one row
      const double* lrow = X + (8 * ib + r) * LD;
      double a0 = 0.0, a1 = 0.0;
      int k = cc & ~7;                         // X_kj = 0 above the diagonal block of column cc
      for (; k + 1 < ncol; k += 2) {
        a0 += lrow[k] * X[k * LD + cc];
        a1 += lrow[k + 1] * X[(k + 1) * LD + cc];
      }
      T[r * LD + cc] = a0 + a1;
    }
    __syncthreads();
    for (int e = tid; e < 8 * ncol; e += 256) {
      const int r = e / ncol, cc = e - r * ncol;
      const double* xr = X + (8 * ib + r) * LD + 8 * ib;   // row r of X_ii
      double a = 0.0;
      B2_UNROLL
      for (int t = 0; t < 8; t++) a += xr[t] * T[t * LD + cc];
      X[(8 * ib + r) * LD + cc] = -a;
    }
    __syncthreads();
  }
  for (int e = tid; e < SB * SB; e += 256) {
    const int i = e % SB, k = e / SB;
    out[e] = X[i * LD + k];
  }
}

// ------------------------------------------------------------------------------------------
// (3b) big fronts: the triangular solves of ONE front are spread over many CTAs, one per
// 64-row chunk (forward) / 64-column block (backward), chained by flags in global memory: a CTA
// only ever waits for CTAs with a smaller VIRTUAL block index of the same launch (the items are
// ordered that way; the virtual index is a ticket taken from an atomic counter when the CTA starts,
// so it is the order in which the CTAs really became resident, whatever the hardware's dispatch
// order): the waits cannot deadlock.  The strip of the panel a CTA needs next is
// prefetched into registers BEFORE it waits, so the chain per block is: flag round trip +
// 64 x 64 mat-vec + in-block substitution by one warp.
// ------------------------------------------------------------------------------------------
// Publishing between the CTAs of one launch: a slot holds the all-ones NaN pattern (the buffer is
// memset to 0xFF at the start of a sweep) until its value is stored; a consumer polls the slot
// itself, so one L2 round trip hands a value over (no separate flag, no fence: an aligned 8-byte
// store is atomic and nothing else is published with it).
__device__ __forceinline__ double poll_value(const double* p) {
#ifdef B2_EMULATE
  // the emulator runs the CTAs one after the other in blockIdx order: the producer must be done
  unsigned long long u;
  memcpy(&u, p, 8);
  if (u == ~0ull) { fprintf(stderr, "k_*_big: wait on a CTA that has not run\n"); abort(); }
  return *p;
#else
  const volatile unsigned long long* q = reinterpret_cast<const volatile unsigned long long*>(p);
  unsigned long long u = *q;
  while (u == ~0ull) u = *q;          // (a __nanosleep back-off was measured: +1 % on the C4 solve, the hand-over is the critical path)
  return __longlong_as_double((long long)u);
#endif
}

__device__ __forceinline__ void publish_value(double* p, double v) {
#ifdef B2_EMULATE
  unsigned long long u;
  memcpy(&u, &v, 8);
  if (u == ~0ull) u = 0x7ff8000000000000ull;
  memcpy(p, &u, 8);
#else
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  if (u == ~0ull) u = 0x7ff8000000000000ull;   // never publish the sentinel itself
  *reinterpret_cast<volatile unsigned long long*>(p) = u;
#endif
}

// item = (front, chunk, gather id, -): chunk c < nblk owns pivot rows
// [64c, min(64c+64, w)); chunk c >= nblk owns the rows [w + 64(c-nblk), ...) below the pivots.
// entries (sb_ent: first / end child row; sb_off: offsets of the child's rel and update vector) list the child update vectors
// that land in the chunk, in child order.
__global__ void __launch_bounds__(256) k_fwd_big(PlanDev P, const int32_t* __restrict__ items, int nitems,
                                                 double* __restrict__ x, double* __restrict__ upd,
                                                 double* __restrict__ ypub, int* __restrict__ ticket) {
  __shared__ int s_vb;
  if (threadIdx.x == 0) s_vb = atomicAdd(ticket, 1);
  __syncthreads();
  const int b = s_vb;
  if (b >= nitems) return;
  const int s = items[4 * b], c = items[4 * b + 1];
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int m = (int)(P.rptr[s + 1] - P.rptr[s]);
  const int nblk = (w + SB - 1) / SB;
  const bool pivot = c < nblk;
  const int i0 = pivot ? c * SB : w + (c - nblk) * SB;
  const int nrow = min(SB, (pivot ? w : m) - i0);
  const double* Lp = P.Lx + P.lptr[s];
  __shared__ double t[SB], yb[SB], red[4][SB];
  __shared__ double Ld[SB * (SB + 1)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = tid & 63, q = tid >> 6;
  // t = own entries of the right-hand side + the children's update vectors: a CSR gather built on
  // the host (row r of the chunk sums upd[sb_src[e]], e in [sb_ptr[g + r], sb_ptr[g + r + 1]), in
  // child order => deterministic), one thread per row, no barriers, loads independent
  if (tid < SB) {
    double acc = (pivot && tid < nrow) ? x[c0 + i0 + tid] : 0.0;
    const int64_t gb = (int64_t)items[4 * b + 2] * (SB + 1);
    const int64_t e1 = P.sb_ptr[gb + tid + 1];
#pragma unroll 4
    for (int64_t e = P.sb_ptr[gb + tid]; e < e1; e++) acc += upd[P.sb_src[e]];
    t[tid] = acc;
  }
  __syncthreads();
  const int li = P.linv_idx[s];              // >= 0: explicit inverses of the diagonal blocks (k_linv)
  if (pivot) {  // own diagonal block (prefetched before any wait): its inverse, or the strict lower part
    if (li >= 0) {
      const double* Xi = P.Linv + (size_t)(li + c) * SB * SB;
      for (int e = tid; e < SB * SB; e += 256) Ld[(e % SB) + (e / SB) * (SB + 1)] = Xi[e];
    } else {
      for (int e = tid; e < SB * SB; e += 256) {
        const int i = e % SB, j = e / SB;
        Ld[i + j * (SB + 1)] = (i < nrow && j < i) ? Lp[(i0 + i) + (size_t)(i0 + j) * m] : 0.0;
      }
    }
  }
  __syncthreads();   // Ld (and t) complete before warp 0 uses them: chunk 0 has no wait in between
  const int nprev = pivot ? c : nblk;
  double lreg[16];
  auto prefetch = [&](int blk) {
    const int kw = min(SB, w - blk * SB);
    B2_UNROLL
    for (int k = 0; k < 16; k++) {
      const int kk = q * 16 + k;
      lreg[k] = (r < nrow && kk < kw) ? Lp[(i0 + r) + (size_t)(blk * SB + kk) * m] : 0.0;
    }
  };
  if (nprev > 0) prefetch(0);
  for (int blk = 0; blk < nprev; blk++) {
    if (tid < SB) yb[tid] = (blk * SB + tid < w) ? poll_value(ypub + c0 + blk * SB + tid) : 0.0;
    __syncthreads();
    double acc = 0.0, acc2 = 0.0;
    B2_UNROLL
    for (int k = 0; k < 16; k += 2) {
      acc += lreg[k] * yb[q * 16 + k];
      acc2 += lreg[k + 1] * yb[q * 16 + k + 1];
    }
    red[q][r] = acc + acc2;
    if (blk + 1 < nprev) prefetch(blk + 1);
    __syncthreads();
    if (tid < SB) t[tid] -= (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
    __syncthreads();
  }
  if (pivot && li >= 0) {
    // y = L_cc^{-1} t as a mat-vec by the whole CTA: thread <-> (row r, quarter q of the columns)
    double acc = 0.0, acc2 = 0.0;
    B2_UNROLL
    for (int k = 0; k < 16; k += 2) {
      acc += Ld[r + (q * 16 + k) * (SB + 1)] * t[q * 16 + k];
      acc2 += Ld[r + (q * 16 + k + 1) * (SB + 1)] * t[q * 16 + k + 1];
    }
    red[q][r] = acc + acc2;
    __syncthreads();
    if (tid < nrow) {
      const double y = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
      publish_value(ypub + c0 + i0 + tid, y);
      x[c0 + i0 + tid] = y / P.dvec[c0 + i0 + tid];
    }
  } else if (pivot) {
    // unit-lower substitution inside the block by warp 0, 8 columns at a time: every lane solves
    // the 8 x 8 triangle redundantly in registers (no communication on the chain), then updates
    // its two rows (lane, lane + 32) below it
    if (warp == 0) {
      double t0 = t[lane], t1 = t[lane + 32];
      for (int kb = 0; kb < nrow; kb += 8) {
        double y8[8];
        B2_UNROLL
        for (int c = 0; c < 8; c++) y8[c] = t[kb + c];
        B2_UNROLL
        for (int c = 1; c < 8; c++)
          B2_UNROLL
          for (int tt = 0; tt < c; tt++) y8[c] -= Ld[(kb + c) + (kb + tt) * (SB + 1)] * y8[tt];
        double u0 = 0.0, u1 = 0.0;
        B2_UNROLL
        for (int c = 0; c < 8; c++) {
          const double* col = Ld + (kb + c) * (SB + 1);
          u0 += col[lane] * y8[c];
          u1 += col[lane + 32] * y8[c];
        }
        __syncwarp();
        // rows inside the 8-block take the solved values, rows below it the update
        B2_UNROLL
        for (int c = 0; c < 8; c++) {
          if (lane == kb + c) t0 = y8[c];
          if (lane + 32 == kb + c) t1 = y8[c];
        }
        if (lane >= kb + 8) t0 -= u0;
        if (lane + 32 >= kb + 8) t1 -= u1;
        t[lane] = t0;
        t[lane + 32] = t1;
        __syncwarp();
      }
    }
    __syncthreads();
    if (tid < nrow) {
      publish_value(ypub + c0 + i0 + tid, t[tid]);
      x[c0 + i0 + tid] = t[tid] / P.dvec[c0 + i0 + tid];
    }
  } else {
    if (tid < nrow) upd[P.uptr[s] + (i0 - w) + tid] = t[tid];
  }
}

// item = (front, column block), blocks of a front in DESCENDING order.
__global__ void __launch_bounds__(256) k_bwd_big(PlanDev P, const int32_t* __restrict__ items, int nitems,
                                                 double* __restrict__ x, double* __restrict__ xpub,
                                                 int* __restrict__ ticket) {
  __shared__ int s_vb;
  if (threadIdx.x == 0) s_vb = atomicAdd(ticket, 1);
  __syncthreads();
  const int b = s_vb;
  if (b >= nitems) return;
  const int s = items[2 * b], c = items[2 * b + 1];
  const int c0 = P.scol[s], w = P.scol[s + 1] - c0;
  const int64_t r0 = P.rptr[s];
  const int m = (int)(P.rptr[s + 1] - r0);
  const int nblk = (w + SB - 1) / SB;
  const int j0 = c * SB, ncol = min(SB, w - j0);
  const double* Lp = P.Lx + P.lptr[s];
  B2_DYN_SMEM(raw);
  double* xb = reinterpret_cast<double*>(raw);        // x of the rows below the pivots (m - w)
  __shared__ double sc[SB], xd[SB];
  __shared__ double Ld[SB * (SB + 1)];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  B2_TICK(54);
  if (tid < SB) sc[tid] = (tid < ncol) ? x[c0 + j0 + tid] : 0.0;
  for (int i = w + tid; i < m; i += 256) xb[i - w] = x[P.rowidx[r0 + i]];
  const int li = P.linv_idx[s];                // >= 0: explicit inverses of the diagonal blocks (k_linv)
  if (li >= 0) {
    const double* Xi = P.Linv + (size_t)(li + c) * SB * SB;
    for (int e = tid; e < SB * SB; e += 256) Ld[(e % SB) + (e / SB) * (SB + 1)] = Xi[e];
  } else {
    for (int e = tid; e < SB * SB; e += 256) {   // own diagonal block: Ld[i][j] = L(j0+i, j0+j), j < i
      const int i = e % SB, j = e / SB;
      Ld[i + j * (SB + 1)] = (i < ncol && j < i) ? Lp[(j0 + i) + (size_t)(j0 + j) * m] : 0.0;
    }
  }
  __syncthreads();
  B2_TICK(55);
  // rows below the pivots: sc[k] -= sum_i L(i, j0+k) xb[i]; every warp takes the columns
  // warp, warp + 8, ... together (one pass over the rows, eight independent sums), lanes over rows
  if (m > w) {
    double acc[8];
    const double* colp[8];
    B2_UNROLL
    for (int kk = 0; kk < 8; kk++) {
      acc[kk] = 0.0;
      colp[kk] = Lp + (size_t)(j0 + min(warp + 8 * kk, ncol - 1)) * m;
    }
    for (int i = w + lane; i < m; i += 32) {
      const double xv = xb[i - w];
      B2_UNROLL
      for (int kk = 0; kk < 8; kk++) acc[kk] += colp[kk][i] * xv;
    }
    B2_UNROLL
    for (int kk = 0; kk < 8; kk++) {
      double v = acc[kk];
      B2_UNROLL
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && warp + 8 * kk < ncol) sc[warp + 8 * kk] -= v;
    }
  }
  B2_TICK(56);
  // later pivot blocks of the same front, as they are published
  double lreg[8][2];
  auto prefetch = [&](int d) {
    const int dw = min(SB, w - d * SB);
    B2_UNROLL
    for (int kk = 0; kk < 8; kk++) {
      const int k = warp + 8 * kk;
      const double* col = Lp + (size_t)(j0 + k) * m + d * SB;
      lreg[kk][0] = (k < ncol && lane < dw) ? col[lane] : 0.0;
      lreg[kk][1] = (k < ncol && lane + 32 < dw) ? col[lane + 32] : 0.0;
    }
  };
  if (c + 1 < nblk) prefetch(nblk - 1);
  for (int d = nblk - 1; d > c; d--) {
    if (tid < SB) xd[tid] = (d * SB + tid < w) ? poll_value(xpub + c0 + d * SB + tid) : 0.0;
    __syncthreads();
    double part[8];
    B2_UNROLL
    for (int kk = 0; kk < 8; kk++) part[kk] = lreg[kk][0] * xd[lane] + lreg[kk][1] * xd[lane + 32];
    if (d - 1 > c) prefetch(d - 1);
    B2_UNROLL
    for (int kk = 0; kk < 8; kk++) {
      double v = part[kk];
      B2_UNROLL
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && warp + 8 * kk < ncol) sc[warp + 8 * kk] -= v;
    }
    __syncthreads();
  }
  __syncthreads();
  B2_TICK(57);
  if (li >= 0) {
    // x = L_cc^{-T} sc as a mat-vec by the whole CTA: thread <-> (column r, quarter q of the rows)
    __shared__ double redb[4][SB];
    const int r = tid & 63, q = tid >> 6;
    double acc = 0.0, acc2 = 0.0;
    B2_UNROLL
    for (int k = 0; k < 16; k += 2) {
      acc += Ld[(q * 16 + k) + r * (SB + 1)] * sc[q * 16 + k];
      acc2 += Ld[(q * 16 + k + 1) + r * (SB + 1)] * sc[q * 16 + k + 1];
    }
    redb[q][r] = acc + acc2;
    __syncthreads();
    if (tid < ncol) {
      const double xv = (redb[0][tid] + redb[1][tid]) + (redb[2][tid] + redb[3][tid]);
      x[c0 + j0 + tid] = xv;
      publish_value(xpub + c0 + j0 + tid, xv);
    }
    B2_TICK(58);
    return;
  }
  // L_cc^T x = sc inside the block by warp 0, 8 columns at a time from the end: every lane solves
  // the 8 x 8 triangle redundantly, then updates its two entries (lane, lane + 32) before it
  if (warp == 0) {
    double s0 = sc[lane], s1 = sc[lane + 32];
    for (int kb = ((ncol - 1) >> 3) << 3; kb >= 0; kb -= 8) {
      double x8[8];
      B2_UNROLL
      for (int c2 = 0; c2 < 8; c2++) x8[c2] = sc[kb + c2];
      B2_UNROLL
      for (int c2 = 6; c2 >= 0; c2--)
        B2_UNROLL
        for (int tt = c2 + 1; tt < 8; tt++) x8[c2] -= Ld[(kb + tt) + (kb + c2) * (SB + 1)] * x8[tt];
      double u0 = 0.0, u1 = 0.0;   // entries j < kb: s_j -= sum_c L(j0+kb+c, j0+j) x_c = Ld[(kb+c) + j*(SB+1)]
      B2_UNROLL
      for (int c2 = 0; c2 < 8; c2++) {
        u0 += Ld[(kb + c2) + lane * (SB + 1)] * x8[c2];
        u1 += Ld[(kb + c2) + (lane + 32) * (SB + 1)] * x8[c2];
      }
      __syncwarp();
      B2_UNROLL
      for (int c2 = 0; c2 < 8; c2++) {
        if (lane == kb + c2) s0 = x8[c2];
        if (lane + 32 == kb + c2) s1 = x8[c2];
      }
      if (lane < kb) s0 -= u0;
      if (lane + 32 < kb) s1 -= u1;
      sc[lane] = s0;
      sc[lane + 32] = s1;
      __syncwarp();
    }
    if (lane < ncol) { x[c0 + j0 + lane] = s0; publish_value(xpub + c0 + j0 + lane, s0); }
    if (lane + 32 < ncol) { x[c0 + j0 + lane + 32] = s1; publish_value(xpub + c0 + j0 + lane + 32, s1); }
  }
  B2_TICK(58);
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
