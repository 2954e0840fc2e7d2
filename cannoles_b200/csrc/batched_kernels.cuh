// batched_kernels.cuh -- one CTA per KKT system, the whole (permuted) lower triangle resident in
// shared memory as a packed dense array; used for batches of small independent systems that share
// one sparsity pattern (BASELINE.json config 5: N = 208 -> 21 736 doubles = 174 KB).
//
// Per instance, fused in ONE kernel (no factor ever leaves the SM unless asked for):
//   assemble   COO values -> packed lower triangle of P K P' (duplicates summed in COO order)
//   factorize  level-phased LEFT-looking supernodal LDL^T without pivoting: for every supernode
//              of a tree level, panel -= L(:,K) D(K) L(J,K)' over its contributing columns K
//              (FP64 tensor-core tiles, mma.sync m8n8k4 = SASS DMMA), then the dense panel
//              factorization; structural zeros stay exact zeros in the packed array
//   inertia    pivot-sign counts (src/solver_types.jl:90-96)
//   solve      forward / diagonal / backward substitution on the shared-memory factor
// Algorithmic bytes per instance: 8 nnz in (+ 8 N rhs) and 8 N out; flops: sum_j (c_j^2 + 3 c_j).
#pragma once
#include "b2_cuda.h"
#include "mma.cuh"

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
  const int32_t* dst_single; // nnz: packed destination of a COO entry that is alone in its slot, else -1
  const int32_t* multi_dst;  // nmulti
  const int32_t* multi_ptr;  // nmulti+1 into multi_coo
  const int32_t* multi_coo;  // COO indices, ascending inside a slot
};

// flags: bit0 = write the packed factor to Lout, bit1 = solve with rhs -> dout (only if the
// inertia is the expected one), bit2 = negate the solution (solve_ldl!'s sign flip),
// bit3 = read the factor from Lout instead of factorizing (solve-only call)
constexpr int BF_STORE = 1, BF_SOLVE = 2, BF_NEGATE = 4, BF_LOAD = 8;

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
  double* Pk = reinterpret_cast<double*>(raw);   // packed lower triangle
  double* xs = Pk + P.npacked;                   // N: solve vector / column scratch
  double* lk = xs + P.N;                         // N
  __shared__ int cnt[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  const int N = P.N;
  if (tid < 4) cnt[tid] = 0;

  B2_TICK(30);
  if (flags & BF_LOAD) {
    const double* src = Lout + (size_t)b * P.npacked;
    for (int i = tid; i < P.npacked; i += NT) Pk[i] = src[i];
    __syncthreads();
  } else {
    // ---------------------------------------------------------------- assemble
    for (int i = tid; i < P.npacked; i += NT) Pk[i] = 0.0;
    __syncthreads();
    const double* v = vals + (size_t)b * P.nnz;
    const int t_rho = P.nnz - P.nvar, t_del = t_rho - P.ncon;
    for (int t = tid; t < P.nnz; t += NT) {
      const int dst = P.dst_single[t];
      if (dst < 0) continue;
      double val = v[t];
      if (rho_over && t >= t_rho) val = rho_over[b];
      else if (delta_over && t >= t_del && t < t_rho) val = -delta_over[b];
      Pk[dst] = 0.0 + val;
    }
    for (int q = tid; q < P.nmulti; q += NT) {
      double acc = 0.0;
      for (int p = P.multi_ptr[q]; p < P.multi_ptr[q + 1]; p++) {
        const int t = P.multi_coo[p];
        double val = v[t];
        if (rho_over && t >= t_rho) val = rho_over[b];
        else if (delta_over && t >= t_del && t < t_rho) val = -delta_over[b];
        acc += val;
      }
      Pk[P.multi_dst[q]] = acc;
    }
    __syncthreads();
    B2_TICK(31);
    // ---------------------------------------------------------------- factorize
    for (int ph = 0; ph < P.nphase; ph++) {
      // (a) left-looking update of every panel of this level, 8 x 8 tiles on the tensor cores
      for (int qs = P.ph_ptr[ph]; qs < P.ph_ptr[ph + 1]; qs++) {
        const int s = P.ph_sn[qs];
        const int nct = P.ct_ptr[s + 1] - P.ct_ptr[s];
        if (nct == 0) continue;
        const int c0 = P.sc0[s], w = P.sc0[s + 1] - c0;
        const int nrb = P.rb_ptr[s + 1] - P.rb_ptr[s];
        const int m = w + nrb;
        const int32_t* rb = P.rb_idx + P.rb_ptr[s];
        const int32_t* ct = P.ct_col + P.ct_ptr[s];
        const int ntj = (w + 7) >> 3, nti = (m + 7) >> 3;
        const int ntiles = ntj * nti;
        for (int tile = warp; tile < ntiles; tile += NW) {
          const int tj = tile / nti, ti = tile - tj * nti;
          if (ti < tj) continue;
          // A fragment row (target row) and B fragment column (target column) of this lane
          const int ar = ti * 8 + (lane >> 2);
          const int grow = ar < w ? c0 + ar : (ar < m ? rb[ar - w] : -1);
          const int bc = tj * 8 + (lane >> 2);
          const int gcol = bc < w ? c0 + bc : -1;
          double acc0 = 0.0, acc1 = 0.0;
          for (int kk = 0; kk < nct; kk += 4) {
            const int ki = kk + (lane & 3);
            double a = 0.0, bb = 0.0;
            if (ki < nct) {
              const int k = ct[ki];
              const int base = P.cbm[k];
              if (grow >= 0) a = Pk[base + grow];
              if (gcol >= 0) bb = Pk[base + gcol] * Pk[base + k];
            }
            dmma_8x8x4(acc0, acc1, a, bb);
          }
          // C fragment: row lane/4, columns 2*(lane%4) + {0,1}
          const int cr = ti * 8 + (lane >> 2);
          const int crow = cr < w ? c0 + cr : (cr < m ? rb[cr - w] : -1);
          const int cc = tj * 8 + (lane & 3) * 2;
          if (crow >= 0) {
            if (cc < w && crow >= c0 + cc) Pk[P.cbm[c0 + cc] + crow] -= acc0;
            if (cc + 1 < w && crow >= c0 + cc + 1) Pk[P.cbm[c0 + cc + 1] + crow] -= acc1;
          }
        }
      }
      __syncthreads();
      B2_TICK(32 + 2 * ph);
      // (b) dense factorization of every panel of this level
      //     width-1 supernodes: all in parallel (one scaling per row); wider ones: column loop
      for (int qs = P.ph_ptr[ph] + warp; qs < P.ph_ptr[ph + 1]; qs += NW) {
        const int s = P.ph_sn[qs];
        const int c0 = P.sc0[s];
        if (P.sc0[s + 1] - c0 != 1) continue;
        const int nrb = P.rb_ptr[s + 1] - P.rb_ptr[s];
        const int32_t* rb = P.rb_idx + P.rb_ptr[s];
        const int base = P.cbm[c0];
        const double dk = Pk[base + c0];
        for (int i = lane; i < nrb; i += 32) Pk[base + rb[i]] /= dk;
      }
      for (int qs = P.ph_ptr[ph]; qs < P.ph_ptr[ph + 1]; qs++) {
        const int s = P.ph_sn[qs];
        const int c0 = P.sc0[s], w = P.sc0[s + 1] - c0;
        if (w == 1) continue;
        const int nrb = P.rb_ptr[s + 1] - P.rb_ptr[s];
        const int m = w + nrb;
        const int32_t* rb = P.rb_idx + P.rb_ptr[s];
        __syncthreads();
        for (int k = 0; k < w; k++) {
          const int gk = c0 + k;
          const int base = P.cbm[gk];
          const double dk = Pk[base + gk];
          // rows below the pivot inside the front: local index t in (k, m)
          for (int t = k + 1 + tid; t < m; t += NT) {
            const int g = t < w ? c0 + t : rb[t - w];
            const double a = Pk[base + g];
            const double l = a / dk;
            xs[t] = a;
            lk[t] = l;
            Pk[base + g] = l;
          }
          __syncthreads();
          for (int j = k + 1 + warp; j < w; j += NW) {
            const double ajk = xs[j];
            const int bj = P.cbm[c0 + j];
            for (int t = j + lane; t < m; t += 32) {
              const int g = t < w ? c0 + t : rb[t - w];
              Pk[bj + g] -= lk[t] * ajk;
            }
          }
          __syncthreads();
        }
      }
      __syncthreads();
      B2_TICK(33 + 2 * ph);
    }
    B2_TICK(40);
    // ---------------------------------------------------------------- inertia
    {
      int pos = 0, zer = 0, neg = 0, brk = 0;
      for (int k = tid; k < N; k += NT) {
        const double d = Pk[P.cbm[k] + k];
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
      if (tid < 4 && counts4) counts4[(size_t)b * 4 + tid] = cnt[tid];
    }
    if (flags & BF_STORE) {
      double* dst = Lout + (size_t)b * P.npacked;
      for (int i = tid; i < P.npacked; i += NT) dst[i] = Pk[i];
    }
    if ((flags & BF_SOLVE) && !(cnt[0] == P.nvar && cnt[1] == 0)) return;  // wrong inertia: no solve
  }
  if (!(flags & BF_SOLVE)) return;
  B2_TICK(41);
  // ------------------------------------------------------------------ solve (factor in smem)
  const double* bvec = rhs + (size_t)b * N;
  for (int k = tid; k < N; k += NT) xs[k] = bvec[P.perm[k]];
  __syncthreads();
  // forward: per level, gather the contributions of finished columns, then the pivot block
  for (int ph = 0; ph < P.nphase; ph++) {
    for (int qs = P.ph_ptr[ph]; qs < P.ph_ptr[ph + 1]; qs++) {
      const int s = P.ph_sn[qs];
      const int nct = P.ct_ptr[s + 1] - P.ct_ptr[s];
      if (nct == 0) continue;
      const int c0 = P.sc0[s], w = P.sc0[s + 1] - c0;
      const int32_t* ct = P.ct_col + P.ct_ptr[s];
      for (int j = tid; j < w; j += NT) {
        const int g = c0 + j;
        double acc = 0.0;
        for (int q = 0; q < nct; q++) {
          const int k = ct[q];
          acc += Pk[P.cbm[k] + g] * xs[k];
        }
        lk[g] = acc;
      }
    }
    __syncthreads();
    for (int qs = P.ph_ptr[ph]; qs < P.ph_ptr[ph + 1]; qs++) {
      const int s = P.ph_sn[qs];
      if (P.ct_ptr[s + 1] == P.ct_ptr[s]) continue;
      const int c0 = P.sc0[s], w = P.sc0[s + 1] - c0;
      for (int j = tid; j < w; j += NT) xs[c0 + j] -= lk[c0 + j];
    }
    __syncthreads();
    // unit-lower pivot blocks: one warp per supernode, columns in order
    for (int qs = P.ph_ptr[ph] + warp; qs < P.ph_ptr[ph + 1]; qs += NW) {
      const int s = P.ph_sn[qs];
      const int c0 = P.sc0[s], w = P.sc0[s + 1] - c0;
      for (int k = 0; k + 1 < w; k++) {
        const double yk = xs[c0 + k];
        const int base = P.cbm[c0 + k];
        for (int i = k + 1 + lane; i < w; i += 32) xs[c0 + i] -= Pk[base + c0 + i] * yk;
        __syncwarp();
      }
    }
    __syncthreads();
  }
  for (int k = tid; k < N; k += NT) xs[k] /= Pk[P.cbm[k] + k];
  __syncthreads();
  // backward: levels in reverse; pivot block first, then nothing else is needed because the
  // rows below a supernode belong to higher levels and are already final
  for (int ph = P.nphase - 1; ph >= 0; ph--) {
    for (int qs = P.ph_ptr[ph] + warp; qs < P.ph_ptr[ph + 1]; qs += NW) {
      const int s = P.ph_sn[qs];
      const int c0 = P.sc0[s], w = P.sc0[s + 1] - c0;
      const int nrb = P.rb_ptr[s + 1] - P.rb_ptr[s];
      const int32_t* rb = P.rb_idx + P.rb_ptr[s];
      // x_k -= sum_{i in rows below} L(i,k) x_i   (one lane per column, rows sequential)
      for (int j = lane; j < w; j += 32) {
        const int base = P.cbm[c0 + j];
        double acc = 0.0;
        for (int i = 0; i < nrb; i++) acc += Pk[base + rb[i]] * xs[rb[i]];
        xs[c0 + j] -= acc;
      }
      __syncwarp();
      for (int k = w - 1; k > 0; k--) {
        const double xk = xs[c0 + k];
        // x_j -= L(k, j) x_k for j < k
        for (int j = lane; j < k; j += 32) xs[c0 + j] -= Pk[P.cbm[c0 + j] + c0 + k] * xk;
        __syncwarp();
      }
    }
    __syncthreads();
  }
  const double sign = (flags & BF_NEGATE) ? -1.0 : 1.0;
  double* dv = dout + (size_t)b * N;
  for (int k = tid; k < N; k += NT) dv[P.perm[k]] = sign * xs[k];
  B2_TICK(42);
}

}  // namespace b2
