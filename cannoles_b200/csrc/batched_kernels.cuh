// batched_kernels.cuh -- one CTA per KKT system, the whole (permuted) lower triangle resident in
// shared memory as a packed dense array; used for batches of small independent systems that share
// one sparsity pattern (BASELINE.json config 5: N = 208 -> 21 736 doubles = 174 KB).
//
// Per instance, fused in ONE kernel (no factor ever leaves the SM unless asked for):
//   assemble   COO values -> packed lower triangle of P K P' (duplicates summed in COO order)
//   factorize  level-phased LEFT-looking supernodal LDL^T without pivoting.  For every supernode
//              of a tree level: (a) panel -= L(:,K) D(K) L(J,K)' over its contributing columns K,
//              8 x 8 x 4 FP64 tensor-core tiles (mma.sync m8n8k4 = SASS DMMA), each warp a strip of
//              up to four row tiles so that the B fragment is reused; (b) the dense panel
//              factorization in 8-column blocks: every thread factors the 8 x 8 diagonal block
//              redundantly in registers (one reciprocal per pivot, no communication on the pivot
//              chain), substitutes its own row, then a rank-8 DMMA update of the remaining columns
//   inertia    pivot-sign counts (src/solver_types.jl:90-96)
//   solve      forward / diagonal / backward substitution on the shared-memory factor, the dense
//              pivot blocks again in 8-column blocks
// Algorithmic bytes per instance: 8 nnz in (+ 8 N rhs) and 8 N out; flops: sum_j (c_j^2 + 3 c_j).
#pragma once
#include "b2_cuda.h"
#include "mma.cuh"
#include "ldlt_packed.cuh"

namespace b2 {

struct BatchPlanDev {
  int N, nnz, nsuper, nphase, npacked;
  int nvar, nequ, ncon;
  int nmulti;                // CSC slots with more than one COO contributor
  const int32_t* perm;       // N: perm[k] = original index of pivot k
  const int32_t* cbm;        // N: packed offset of column k minus k  (L(i,k) = P[cbm[k] + i])
  const int32_t* sc0;        // nsuper+1: first column of each supernode
  const int32_t* rb_ptr;     // nsuper+1: rows below the pivot block (permuted indices, ascending)
  const int32_t* rb_idx;
  const int32_t* ct_ptr;     // nsuper+1: contributing columns (k < sc0[s]), ascending
  const int32_t* ct_col;
  const int32_t* ph_ptr;     // nphase+1: supernodes grouped by tree level
  const int32_t* ph_sn;
  // the same per level without the supernodes the sequential loops would only skip: 80 dependent
  // plan reads per loop to skip the 79 leaves of config 5 cost 13 k cycles per loop (measured)
  const int32_t* phw_ptr;    // nphase+1: supernodes that are not (width 1, no contributions): factorization, forward
  const int32_t* phw_sn;
  const int32_t* phb_ptr;    // nphase+1: supernodes of width > 1: backward
  const int32_t* phb_sn;
  // width-1 supernodes per level as FLAT lists (no per-supernode chain of dependent plan reads): the
  // supernodes without contributing columns (leaves of the factorization) first
  const int32_t* lw_ptr;     // nphase+1 into lw_meta
  const int32_t* lw_nleaf;   // nphase: how many of the level's width-1 supernodes are factorization leaves
  const int4* lw_meta;       // (column k, first entry, entries = rows below, cbm[k]); one sentinel at the end
  const int2* lf_ent;        // per row below: (packed index of L(row, k), k | row << 16)
  const int32_t* dst_single; // nnz: packed destination of a COO entry that is alone in its slot, else -1
  const int32_t* multi_dst;  // nmulti
  const int32_t* multi_ptr;  // nmulti+1 into multi_coo
  const int32_t* multi_coo;  // COO indices, ascending inside a slot
};

// shared memory of k_batched, in bytes, for a system of order N
__host__ __device__ inline size_t batched_smem_bytes(int N, int64_t npacked) {
  return ((size_t)npacked + 11 * (size_t)N) * sizeof(double) + 3 * (size_t)N * sizeof(int32_t);
}

// flags: bit0 = write the packed factor to Lout, bit1 = solve with rhs -> dout (only if the
// inertia is the expected one), bit2 = negate the solution (solve_ldl!'s sign flip),
// bit3 = read the factor from Lout instead of factorizing (solve-only call)
constexpr int BF_STORE = 1, BF_SOLVE = 2, BF_NEGATE = 4, BF_LOAD = 8;
#ifndef B2_BATCH_NT
#define B2_BATCH_NT 256
#endif
#ifndef B2_BATCH_GRP
#define B2_BATCH_GRP 4
#endif
constexpr int BATCH_GRP = B2_BATCH_GRP;  // 8-column steps per group of the dense panel factorization (see k_batched (b))
constexpr int BATCH_NT = B2_BATCH_NT;   // threads per instance (one CTA per SM: shared memory holds one system)

// C(t, j) -= sum_{q < K} A(t, q) B(j, q) for local columns j in [jbeg, jend) and local rows t in
// [j, m) of one supernode panel, in 8 x 8 tiles; each warp takes strips of up to STRIP row tiles.
//   A(t, q) = Pk[abase[q] + rowmap[t]]
//   B(j, q) = FROM_W ? Wd[j * 8 + q] : Pk[abase[q] + c0 + j] * dd[q]
//   C(t, j) = Pk[cbm[c0 + j] + rowmap[t]]
// STRIP row tiles per task: 4 amortises the B fragment over more DMMA; 2 gives a one-tile-column
// update (the in-group updates of the dense panel) enough tasks for all the warps.
template <int NW, bool FROM_W, int STRIP = 4>
__device__ __forceinline__ void panel_update(double* Pk, const int32_t* cbm, const int32_t* rowmap,
                                             const int32_t* abase, const double* dd, const double* Wd,
                                             int c0, int jbeg, int jend, int m, int K, int warp, int lane) {
  const int g = lane >> 2, t4 = lane & 3;
  const int ntj = (jend - jbeg + 7) >> 3, nti = (m - jbeg + 7) >> 3;
  int task = 0;
  for (int tj = 0; tj < ntj; tj++) {
    for (int ti0 = tj; ti0 < nti; ti0 += STRIP, task++) {
      if (task % NW != warp) continue;
      double acc[STRIP][2];
      B2_UNROLL
      for (int a = 0; a < STRIP; a++) { acc[a][0] = 0.0; acc[a][1] = 0.0; }
      const int jb = jbeg + tj * 8 + g;                 // B-fragment column (local)
      int ra[STRIP];
      B2_UNROLL
      for (int a = 0; a < STRIP; a++) {
        const int tr = jbeg + (ti0 + a) * 8 + g;        // A-fragment row (local)
        ra[a] = (ti0 + a < nti && tr < m) ? rowmap[tr] : -1;
      }
#pragma unroll 4
      for (int q0 = 0; q0 < K; q0 += 4) {
        const int q = q0 + t4;
        double bv = 0.0, av[STRIP];
        B2_UNROLL
        for (int a = 0; a < STRIP; a++) av[a] = 0.0;
        if (q < K) {
          const int base = abase[q];
          if (jb < jend) bv = FROM_W ? Wd[jb * 8 + q] : Pk[base + c0 + jb] * dd[q];
          B2_UNROLL
          for (int a = 0; a < STRIP; a++)
            if (ra[a] >= 0) av[a] = Pk[base + ra[a]];
        }
        B2_UNROLL
        for (int a = 0; a < STRIP; a++) dmma_8x8x4(acc[a][0], acc[a][1], av[a], bv);
      }
      const int cj = jbeg + tj * 8 + 2 * t4;            // C-fragment columns cj, cj + 1 (local)
      B2_UNROLL
      for (int a = 0; a < STRIP; a++) {
        const int tr = jbeg + (ti0 + a) * 8 + g;
        if (ti0 + a >= nti || tr >= m) continue;
        const int gr = rowmap[tr];
        if (cj < jend && tr >= cj) Pk[cbm[c0 + cj] + gr] -= acc[a][0];
        if (cj + 1 < jend && tr >= cj + 1) Pk[cbm[c0 + cj + 1] + gr] -= acc[a][1];
      }
    }
  }
}

// Pivot-free LDL^T of an 8 x 8 block held by EVERY thread in registers (g[c][t], t <= c):
// afterwards g[c][t] (t < c) = L, g[c][c] = d_c, rd[c] = 1 / d_c.
__device__ __forceinline__ void ldlt8_regs(double (&g)[8][8], double (&rd)[8]) {
  B2_UNROLL
  for (int c = 0; c < 8; c++) {
    rd[c] = rcp_nr(g[c][c]);
    // the NEXT pivot first and with one operation after the reciprocal (its square is formed
    // while the reciprocal is in flight): this element is the critical path of the block
    if (c + 1 < 8) g[c + 1][c + 1] = __fma_rn(-(g[c + 1][c] * g[c + 1][c]), rd[c], g[c + 1][c + 1]);
    B2_UNROLL
    for (int r = c + 1; r < 8; r++) {
      const double lrc = g[r][c] * rd[c];
      B2_UNROLL
      for (int t = c + 1; t <= r; t++)
        if (!(r == c + 1 && t == c + 1)) g[r][t] -= lrc * g[t][c];   // column c still unscaled
    }
    B2_UNROLL
    for (int r = c + 1; r < 8; r++) g[r][c] *= rd[c];
  }
}

// The work of ONE instance by one CTA (see the header of this file); `raw` = batched_smem_bytes of
// shared memory, `cnt` = 4 ints of shared memory that hold the pivot-sign counts on return
// (valid after the barrier the function ends its factorization part with).  v = the instance's COO
// values, rho / -delta override the trailing segments when has_rho / has_del, bvec / dv = its
// right-hand side / solution, Lb = its stored factor, c4 = its four output counters (or NULL).
// Shared by k_batched and by the device-resident solver loop (nls_kernels.cuh).
template <int NT>
__device__ __forceinline__ void batched_instance(const BatchPlanDev& P, unsigned char* raw, int* cnt,
                                                 const double* v, bool has_rho, double rho_b,
                                                 bool has_del, double mdel_b, double eig_tol,
                                                 long long* c4, double* Lb,
                                                 const double* bvec, double* dv, int flags) {
  // (no __restrict__ here: the solver loop writes v / bvec from the same kernel, and bvec / dv may be
  // shared memory; k_batched's own parameters keep the qualifier)
  const int N = P.N;
  double* Pk = reinterpret_cast<double*>(raw);   // packed lower triangle
  double* xs = Pk + P.npacked;                   // N: solve vector
  double* lk = xs + N;                           // N: scratch
  double* Wd = lk + N;                           // N x 8: W = L D of the current 8-column block
  double* dd = Wd + 8 * N;                       // N: pivots of the contributing columns
  int32_t* abase = reinterpret_cast<int32_t*>(dd + N);   // N: packed base of the contributing columns
  int32_t* rowmap = abase + N;                   // N: local row -> permuted index
  int32_t* cbm = rowmap + N;                     // N: copy of P.cbm
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  if (tid < 4) cnt[tid] = 0;
  for (int k = tid; k < N; k += NT) cbm[k] = P.cbm[k];
  B2_TICK(30);

  if (flags & BF_LOAD) {
    const double* src = Lb;
    for (int i = tid; i < P.npacked; i += NT) Pk[i] = src[i];
    __syncthreads();
  } else {
    // ---------------------------------------------------------------- assemble
    for (int i = tid; i < P.npacked; i += NT) Pk[i] = 0.0;
    __syncthreads();
    B2_TICK(20);
    const int t_rho = P.nnz - P.nvar, t_del = t_rho - P.ncon;
    const bool over = has_rho || has_del;
    constexpr int U = 8;   // loads of U entries in flight per thread (16: no faster, measured)
    for (int t0 = tid; t0 < P.nnz; t0 += NT * U) {
      int dst[U];
      double val[U];
      B2_UNROLL
      for (int u = 0; u < U; u++) {
        const int t = t0 + u * NT;
        dst[u] = t < P.nnz ? P.dst_single[t] : -1;
        val[u] = t < P.nnz ? v[t] : 0.0;
      }
      B2_UNROLL
      for (int u = 0; u < U; u++) {
        const int t = t0 + u * NT;
        if (dst[u] < 0) continue;
        double x = val[u];
        if (over) {
          if (has_rho && t >= t_rho) x = rho_b;
          else if (has_del && t >= t_del && t < t_rho) x = mdel_b;
        }
        Pk[dst[u]] = 0.0 + x;
      }
    }
    B2_TICK(21);
    // slots with several COO contributors (every entry of the (1,1) block has two or three: residual
    // Hessian, constraint Hessian, rho): summed in COO order from +0.0 like set_vals!, but the loads of
    // SU slots x 3 contributors are issued together -- three dependent rounds of loads per pass instead
    // of three per slot (17 k -> 6 k cycles on config 5)
    {
      constexpr int SU = 4, MC = 3;
      auto value_of = [&](int t) -> double {
        double x = v[t];
        if (over) {
          if (has_rho && t >= t_rho) x = rho_b;
          else if (has_del && t >= t_del && t < t_rho) x = mdel_b;
        }
        return x;
      };
      for (int q0 = tid; q0 < P.nmulti; q0 += NT * SU) {
        int p0[SU], p1[SU], tt[SU][MC];
        double xv[SU][MC];
        B2_UNROLL
        for (int u = 0; u < SU; u++) {
          const int q = q0 + u * NT;
          p0[u] = q < P.nmulti ? P.multi_ptr[q] : 0;
          p1[u] = q < P.nmulti ? P.multi_ptr[q + 1] : 0;
        }
        B2_UNROLL
        for (int u = 0; u < SU; u++)
          B2_UNROLL
          for (int k = 0; k < MC; k++) tt[u][k] = p0[u] + k < p1[u] ? P.multi_coo[p0[u] + k] : -1;
        B2_UNROLL
        for (int u = 0; u < SU; u++)
          B2_UNROLL
          for (int k = 0; k < MC; k++) xv[u][k] = tt[u][k] >= 0 ? value_of(tt[u][k]) : 0.0;
        B2_UNROLL
        for (int u = 0; u < SU; u++) {
          const int q = q0 + u * NT;
          if (q >= P.nmulti) continue;
          double acc = 0.0;
          B2_UNROLL
          for (int k = 0; k < MC; k++)
            if (tt[u][k] >= 0) acc += xv[u][k];
          for (int p = p0[u] + MC; p < p1[u]; p++) acc += value_of(P.multi_coo[p]);
          Pk[P.multi_dst[q]] = acc;
        }
      }
    }
    __syncthreads();
    B2_TICK(31);
    // ---------------------------------------------------------------- factorize
    for (int ph = 0; ph < P.nphase; ph++) {
      // width-1 supernodes without contributions (the leaves): reciprocal pivots, then ONE flat pass over
      // all their entries (a warp per leaf paid a chain of four dependent plan reads per leaf)
      {
        const int a0 = P.lw_ptr[ph], nl = P.lw_nleaf[ph];
        if (nl > 0) {
          for (int q = tid; q < nl; q += NT) {
            const int4 mt = P.lw_meta[a0 + q];
            lk[mt.x] = rcp_nr(Pk[mt.w + mt.x]);
          }
          const int e0 = P.lw_meta[a0].y, e1 = P.lw_meta[a0 + nl].y;
          __syncthreads();
          for (int e = e0 + tid; e < e1; e += NT) {
            const int2 en = P.lf_ent[e];
            Pk[en.x] *= lk[en.y & 0xffff];
          }
        }
      }
      __syncthreads();
      B2_TICK(32 + 2 * ph);
      // every other supernode of the level, one after the other, the whole CTA on each
      for (int qs = P.phw_ptr[ph]; qs < P.phw_ptr[ph + 1]; qs++) {
        const int s = P.phw_sn[qs];
        const int c0 = P.sc0[s], w = P.sc0[s + 1] - c0;
        const int nct = P.ct_ptr[s + 1] - P.ct_ptr[s];
        const int nrb = P.rb_ptr[s + 1] - P.rb_ptr[s];
        const int m = w + nrb;
        const int32_t* rb = P.rb_idx + P.rb_ptr[s];
        const int32_t* ct = P.ct_col + P.ct_ptr[s];
        for (int t = tid; t < m; t += NT) rowmap[t] = t < w ? c0 + t : rb[t - w];
        for (int q = tid; q < nct; q += NT) {
          const int k = ct[q];
          abase[q] = cbm[k];
          dd[q] = Pk[cbm[k] + k];
        }
        __syncthreads();
#ifdef B2_TIMING
        long long tq = clock64();
        if (tid == 0 && blockIdx.x == 0) { b2_dbg[50] = 0; b2_dbg[51] = 0; b2_dbg[52] = 0; b2_dbg[53] = 0; }
#endif
        // (a) left-looking update of the panel with its contributing columns
        if (nct > 0)
          panel_update<NW, false>(Pk, cbm, rowmap, abase, dd, Wd, c0, 0, w, m, nct, warp, lane);
        __syncthreads();
        B2_ACC(50, tq);
        // (b) dense factorization of the panel in blocks of up to 64 pivot columns: pivot chain on one
        // warp (two pivots per step, look-ahead), substitution and trailing update on the others
#ifdef B2_TIMING
        tq = clock64();
#endif
        for (int jb = 0; jb < w; jb += 64) {
          PackedAcc S{Pk, cbm, rowmap, c0, jb};
          cta_ldlt_packed<NT>(S, min(64, w - jb), w - jb, m - jb, Wd, lk);
          __syncthreads();
        }
        B2_ACC(51, tq);
      }
      B2_TICK(33 + 2 * ph);
    }
    B2_TICK(40);
    // ---------------------------------------------------------------- inertia
    {
      int pos = 0, zer = 0, neg = 0, brk = 0;
      for (int k = tid; k < N; k += NT) {
        const double d = Pk[cbm[k] + k];
        pos += d > eig_tol;
        zer += fabs(d) <= eig_tol;
        neg += d < -eig_tol;
        brk += d == 0.0;
      }
      if (pos) atomicAdd(&cnt[0], pos);
      if (zer) atomicAdd(&cnt[1], zer);
      if (neg) atomicAdd(&cnt[2], neg);
      if (brk) atomicAdd(&cnt[3], brk);
      __syncthreads();
      if (tid < 4 && c4) c4[tid] = cnt[tid];
    }
    if (flags & BF_STORE) {
      double* dst = Lb;
      for (int i = tid; i < P.npacked; i += NT) dst[i] = Pk[i];
    }
    if ((flags & BF_SOLVE) && !(cnt[0] == P.nvar && cnt[1] == 0)) return;  // wrong inertia: no solve
  }
  if (!(flags & BF_SOLVE)) return;
  B2_TICK(41);
  // ------------------------------------------------------------------ solve (factor in smem)
  for (int k = tid; k < N; k += NT) xs[k] = bvec[P.perm[k]];
  __syncthreads();
  // forward: L y = b, level by level; a supernode first gathers the contributions of the
  // finished columns it depends on, then solves its own unit-lower pivot block
  for (int ph = 0; ph < P.nphase; ph++) {
    for (int qs = P.phw_ptr[ph]; qs < P.phw_ptr[ph + 1]; qs++) {   // (a leaf has nothing to do going forward)
      const int s = P.phw_sn[qs];
      const int c0 = P.sc0[s], w = P.sc0[s + 1] - c0;
      const int nct = P.ct_ptr[s + 1] - P.ct_ptr[s];
      const int32_t* ct = P.ct_col + P.ct_ptr[s];
      for (int q = tid; q < nct; q += NT) {
        const int k = ct[q];
        abase[q] = cbm[k];
        lk[q] = xs[k];
      }
      __syncthreads();
      if (tid < w && nct > 0) {
        const int gr = c0 + tid;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int q = 0;
        for (; q + 3 < nct; q += 4) {
          a0 += Pk[abase[q] + gr] * lk[q];
          a1 += Pk[abase[q + 1] + gr] * lk[q + 1];
          a2 += Pk[abase[q + 2] + gr] * lk[q + 2];
          a3 += Pk[abase[q + 3] + gr] * lk[q + 3];
        }
        for (; q < nct; q++) a0 += Pk[abase[q] + gr] * lk[q];
        xs[gr] -= (a0 + a1) + (a2 + a3);
      }
      __syncthreads();
      for (int kb = 0; kb + 1 < w; kb += 8) {            // unit-lower pivot block, 8 columns at a time
        const int pw = min(8, w - kb);
        // everything that does not depend on the solved values first, in registers: column bases, the
        // 8 x 8 triangle, this thread's row of the panel
        int cb8[8];
        B2_UNROLL
        for (int c = 0; c < 8; c++) cb8[c] = c < pw ? cbm[c0 + kb + c] : 0;
        double l8[8][8], lrow[8];
        B2_UNROLL
        for (int c = 1; c < 8; c++)
          B2_UNROLL
          for (int t = 0; t < c; t++) l8[c][t] = c < pw ? Pk[cb8[t] + c0 + kb + c] : 0.0;
        const int tr = kb + 8 + tid;
        B2_UNROLL
        for (int c = 0; c < 8; c++) lrow[c] = (tr < w && c < pw) ? Pk[cb8[c] + c0 + tr] : 0.0;
        double y[8];
        B2_UNROLL
        for (int c = 0; c < 8; c++) y[c] = c < pw ? xs[c0 + kb + c] : 0.0;
        B2_UNROLL
        for (int c = 1; c < 8; c++)
          B2_UNROLL
          for (int t = 0; t < c; t++) y[c] -= l8[c][t] * y[t];
        double acc = 0.0;
        B2_UNROLL
        for (int c = 0; c < 8; c++) acc += lrow[c] * y[c];
        __syncthreads();                                 // everybody has read xs[c0 + kb ..]
        if (tr < w) xs[c0 + tr] -= acc;
        if (tid == 0) {
          B2_UNROLL
          for (int c = 0; c < 8; c++)
            if (c < pw) xs[c0 + kb + c] = y[c];
        }
        __syncthreads();
      }
    }
  }
  B2_TICK(43);
  for (int k = tid; k < N; k += NT) xs[k] /= Pk[cbm[k] + k];
  __syncthreads();
  B2_TICK(44);
  // backward: L' x = z, levels in reverse; the rows below a supernode belong to higher levels
  // and are already final
  for (int ph = P.nphase - 1; ph >= 0; ph--) {
    // rows below: width-1 supernodes, eight lanes each (flat lists, fixed-order shuffle tree)
    {
      const int a0 = P.lw_ptr[ph], nw1 = P.lw_ptr[ph + 1] - a0;
      for (int qb = 0; qb < nw1; qb += NT >> 3) {
        const int q = qb + (tid >> 3);
        double acc = 0.0;
        int col = -1;
        if (q < nw1) {
          const int4 mt = P.lw_meta[a0 + q];
          col = mt.x;
          for (int i = tid & 7; i < mt.z; i += 8) {
            const int2 en = P.lf_ent[mt.y + i];
            acc += Pk[en.x] * xs[en.y >> 16];
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (col >= 0 && (tid & 7) == 0) xs[col] -= acc;
      }
    }
    __syncthreads();
    B2_TICK(46 + (ph & 1));
    for (int qs = P.phb_ptr[ph]; qs < P.phb_ptr[ph + 1]; qs++) {
      const int s = P.phb_sn[qs];
      const int c0 = P.sc0[s], w = P.sc0[s + 1] - c0;
      const int nrb = P.rb_ptr[s + 1] - P.rb_ptr[s];
      const int32_t* rb = P.rb_idx + P.rb_ptr[s];
      if (tid < w && nrb > 0) {                          // one thread per column over the rows below
        const int base = cbm[c0 + tid];
        double a0 = 0.0, a1 = 0.0;
        int i = 0;
        for (; i + 1 < nrb; i += 2) {
          a0 += Pk[base + rb[i]] * xs[rb[i]];
          a1 += Pk[base + rb[i + 1]] * xs[rb[i + 1]];
        }
        if (i < nrb) a0 += Pk[base + rb[i]] * xs[rb[i]];
        xs[c0 + tid] -= a0 + a1;
      }
      __syncthreads();
      for (int kb = ((w - 1) >> 3) << 3; kb >= 0; kb -= 8) {   // L11' x = z, 8 columns at a time from the end
        const int pw = min(8, w - kb);
        // operands that do not depend on the solved values first, in registers
        int cb8[8];
        B2_UNROLL
        for (int c = 0; c < 8; c++) cb8[c] = c < pw ? cbm[c0 + kb + c] : 0;
        double l8[8][8], lcol[8];
        B2_UNROLL
        for (int c = 6; c >= 0; c--)
          B2_UNROLL
          for (int t = c + 1; t < 8; t++) l8[c][t] = t < pw ? Pk[cb8[c] + c0 + kb + t] : 0.0;
        {
          const int base = tid < kb ? cbm[c0 + tid] + c0 + kb : 0;
          B2_UNROLL
          for (int c = 0; c < 8; c++) lcol[c] = (tid < kb && c < pw) ? Pk[base + c] : 0.0;   // L(kb + c, tid)
        }
        double x8[8];
        B2_UNROLL
        for (int c = 0; c < 8; c++) x8[c] = c < pw ? xs[c0 + kb + c] : 0.0;
        B2_UNROLL
        for (int c = 6; c >= 0; c--)
          B2_UNROLL
          for (int t = c + 1; t < 8; t++) x8[c] -= l8[c][t] * x8[t];
        double acc = 0.0;                                 // earlier entries j < kb: z_j -= sum_c L(kb+c, j) x_c
        B2_UNROLL
        for (int c = 0; c < 8; c++) acc += lcol[c] * x8[c];
        __syncthreads();
        if (tid < kb) xs[c0 + tid] -= acc;
        if (tid == 0) {
          B2_UNROLL
          for (int c = 0; c < 8; c++)
            if (c < pw) xs[c0 + kb + c] = x8[c];
        }
        __syncthreads();
      }
    }
  }
  B2_TICK(45);
  const double sign = (flags & BF_NEGATE) ? -1.0 : 1.0;
  for (int k = tid; k < N; k += NT) dv[P.perm[k]] = sign * xs[k];
  B2_TICK(42);
}

template <int NT>
__global__ void __launch_bounds__(NT) k_batched(BatchPlanDev P, int batch, const double* __restrict__ vals,
                                                const double* __restrict__ rho_over,
                                                const double* __restrict__ delta_over,
                                                const uint8_t* __restrict__ active, double eig_tol,
                                                long long* __restrict__ counts4, double* __restrict__ Lout,
                                                const double* __restrict__ rhs, double* __restrict__ dout,
                                                int flags) {
  const int b = blockIdx.x;
  if (b >= batch) return;
  if (active && !active[b]) return;
  B2_DYN_SMEM(raw);
  __shared__ int cnt[4];
  batched_instance<NT>(P, raw, cnt, vals + (size_t)b * P.nnz, rho_over != nullptr, rho_over ? rho_over[b] : 0.0,
                       delta_over != nullptr, delta_over ? -delta_over[b] : 0.0, eig_tol,
                       counts4 ? counts4 + (size_t)b * 4 : nullptr, Lout ? Lout + (size_t)b * P.npacked : nullptr,
                       rhs ? rhs + (size_t)b * P.N : nullptr, dout ? dout + (size_t)b * P.N : nullptr, flags);
}

}  // namespace b2
