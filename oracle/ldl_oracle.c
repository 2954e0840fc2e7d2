/*
 * ldl_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * Restates, in plain C, what the reference's `:ldlfactorizations` backend does for the
 * KKT factor/solve path of CaNNOLeS.jl:
 *
 *   reference/src/solver_types.jl:45-51   LDLFactStruct holder            -> orc_t
 *   reference/src/solver_types.jl:61-65   ctor: sparse(cols,rows,..), triu, ldl_analyze -> orc_analyze
 *   reference/src/solver_types.jl:53-59   set_vals! (zero nzval, += per triplet)        -> orc_set_vals
 *   reference/src/solver_types.jl:79-98   try_to_factorize (factor + inertia loop)      -> orc_try_to_factorize
 *   reference/src/solver_types.jl:69-77   solve_ldl! (ldiv! then negate)                -> orc_solve_ldl
 *
 * The arithmetic itself lives in un-vendored dependencies of the reference
 * (LDLFactorizations.jl 0.10.x, reference/Project.toml:21, and AMD.jl / SuiteSparse AMD).
 * Both are restated from their published algorithms:
 *   - ordering: approximate minimum degree (Amestoy, Davis, Duff, SIMAX 17(4) 1996 /
 *     ACM TOMS 30(3) 2004): quotient graph, approximate external degree, aggressive
 *     element absorption, mass elimination, hashed supervariable detection, dense-row
 *     deferral (threshold max(16, 10 sqrt n)), assembly-tree postorder;
 *   - factorization: Davis' up-looking, pivot-free sparse LDL^T (ACM TOMS 31(4) 2005,
 *     "Algorithm 849"), which LDLFactorizations.jl translates: elimination tree + column
 *     counts (symbolic), then for every column k the sparse triangular solve over the
 *     etree reach; failure on an exact zero pivot; no dynamic regularisation;
 *   - solve: P, L, D, L^T, P^T.
 *
 * PARITY STATUS: the reference cannot run here (no Julia) and its tests pin no value at the
 * factorization boundary, so parity is pinned END-TO-END ONLY (known solutions in
 * reference/test/runtests.jl and the hand-derived first KKT system of MGH01CON, SURVEY
 * App. D); the tie-breaking of SuiteSparse AMD is "parity unpinned".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>

typedef int64_t i64;

#define FLIP(i) (-(i) - 2)

/* ------------------------------------------------------------------------------------ */
/* Approximate minimum degree ordering on the pattern of a symmetric matrix.             */
/* Input: CSC pattern (Ap, Ai) of the FULL symmetric matrix, diagonal entries ignored.    */
/* Output: perm[k] = original index of the k-th pivot.                                    */
/* ------------------------------------------------------------------------------------ */

static i64 amd_clear_marks(i64 mark, i64 lemax, i64 *w, i64 n) {
  if (mark < 2 || (mark + lemax < 0)) {
    for (i64 k = 0; k < n; k++)
      if (w[k] != 0) w[k] = 1;
    mark = 2;
  }
  return mark;
}

/* non-recursive DFS of the assembly tree rooted at j; children lists in head/next */
static i64 amd_tree_dfs(i64 j, i64 k, i64 *head, const i64 *next, i64 *post, i64 *stack) {
  i64 top = 0;
  stack[0] = j;
  while (top >= 0) {
    i64 p = stack[top];
    i64 i = head[p];
    if (i == -1) {
      top--;
      post[k++] = p;
    } else {
      head[p] = next[i];
      stack[++top] = i;
    }
  }
  return k;
}

int orc_amd(i64 n, const i64 *Ap, const i64 *Ai, i64 *perm) {
  if (n <= 0) return 0;
  /* ---- build C = pattern of A + A' without the diagonal, with elbow room ---- */
  i64 *cnt = (i64 *)calloc((size_t)n + 1, sizeof(i64));
  /* Input is assumed structurally symmetric already (we build it so); just drop the diagonal. */
  i64 cnz = 0;
  for (i64 j = 0; j < n; j++)
    for (i64 p = Ap[j]; p < Ap[j + 1]; p++)
      if (Ai[p] != j) { cnt[j]++; cnz++; }
  i64 nzmax = cnz + cnz / 5 + 2 * n + 16;
  i64 *Cp = (i64 *)malloc(((size_t)n + 1) * sizeof(i64));
  i64 *Ci = (i64 *)malloc((size_t)nzmax * sizeof(i64));
  i64 *W = (i64 *)malloc(8 * ((size_t)n + 1) * sizeof(i64));
  if (!cnt || !Cp || !Ci || !W) { free(cnt); free(Cp); free(Ci); free(W); return -1; }
  i64 *len = W, *nv = W + (n + 1), *next = W + 2 * (n + 1), *head = W + 3 * (n + 1),
      *elen = W + 4 * (n + 1), *degree = W + 5 * (n + 1), *w = W + 6 * (n + 1),
      *hhead = W + 7 * (n + 1);
  i64 *last;
  i64 *lastbuf = (i64 *)malloc(((size_t)n + 1) * sizeof(i64));
  if (!lastbuf) { free(cnt); free(Cp); free(Ci); free(W); return -1; }
  last = lastbuf;
  {
    i64 q = 0;
    for (i64 j = 0; j < n; j++) {
      Cp[j] = q;
      for (i64 p = Ap[j]; p < Ap[j + 1]; p++)
        if (Ai[p] != j) Ci[q++] = Ai[p];
    }
    Cp[n] = q;
  }
  free(cnt);

  i64 dense = (i64)(10.0 * sqrt((double)n));
  if (dense < 16) dense = 16;
  if (dense > n - 2) dense = n - 2;

  for (i64 k = 0; k < n; k++) len[k] = Cp[k + 1] - Cp[k];
  len[n] = 0;
  for (i64 i = 0; i <= n; i++) {
    head[i] = -1; last[i] = -1; next[i] = -1; hhead[i] = -1;
    nv[i] = 1; w[i] = 1; elen[i] = 0; degree[i] = len[i];
  }
  i64 mark = amd_clear_marks(0, 0, w, n);
  elen[n] = -2; Cp[n] = -1; w[n] = 0;
  i64 nel = 0, mindeg = 0, lemax = 0;

  /* ---- initial degree lists ---- */
  for (i64 i = 0; i < n; i++) {
    i64 d = degree[i];
    if (d == 0) {               /* isolated variable: order it now */
      elen[i] = -2; nel++; Cp[i] = -1; w[i] = 0;
    } else if (d > dense) {     /* dense variable: defer to the very end */
      nv[i] = 0; elen[i] = -1; nel++; Cp[i] = FLIP(n); nv[n]++;
    } else {
      if (head[d] != -1) last[head[d]] = i;
      next[i] = head[d];
      head[d] = i;
    }
  }

  while (nel < n) {
    /* ---- pick a variable of (approximately) minimum degree ---- */
    i64 k = -1;
    for (; mindeg < n && (k = head[mindeg]) == -1; mindeg++) ;
    if (next[k] != -1) last[next[k]] = -1;
    head[mindeg] = next[k];
    i64 elenk = elen[k], nvk = nv[k];
    nel += nvk;

    /* ---- compact the workspace if the new element might not fit ---- */
    if (elenk > 0 && cnz + mindeg >= nzmax) {
      for (i64 j = 0; j < n; j++) {
        i64 p = Cp[j];
        if (p >= 0) { Cp[j] = Ci[p]; Ci[p] = FLIP(j); }
      }
      i64 q = 0;
      for (i64 p = 0; p < cnz;) {
        i64 j = FLIP(Ci[p++]);
        if (j >= 0) {
          Ci[q] = Cp[j];
          Cp[j] = q++;
          for (i64 t = 0; t < len[j] - 1; t++) Ci[q++] = Ci[p++];
        }
      }
      cnz = q;
    }

    /* ---- form the new element Lk from k's variable list and its adjacent elements ---- */
    i64 dk = 0;
    nv[k] = -nvk;
    i64 p = Cp[k];
    i64 pk1 = (elenk == 0) ? p : cnz;
    i64 pk2 = pk1;
    for (i64 k1 = 1; k1 <= elenk + 1; k1++) {
      i64 e, pj, ln;
      if (k1 > elenk) { e = k; pj = p; ln = len[k] - elenk; }
      else { e = Ci[p++]; pj = Cp[e]; ln = len[e]; }
      for (i64 k2 = 1; k2 <= ln; k2++) {
        i64 i = Ci[pj++];
        i64 nvi = nv[i];
        if (nvi <= 0) continue;
        dk += nvi;
        nv[i] = -nvi;
        Ci[pk2++] = i;
        if (next[i] != -1) last[next[i]] = last[i];
        if (last[i] != -1) next[last[i]] = next[i];
        else head[degree[i]] = next[i];
      }
      if (e != k) { Cp[e] = FLIP(k); w[e] = 0; }
    }
    if (elenk != 0) cnz = pk2;
    degree[k] = dk;
    Cp[k] = pk1;
    len[k] = pk2 - pk1;
    elen[k] = -2;

    /* ---- scan 1: |Le \ Lk| for every element e adjacent to a variable of Lk ---- */
    mark = amd_clear_marks(mark, lemax, w, n);
    for (i64 pk = pk1; pk < pk2; pk++) {
      i64 i = Ci[pk];
      i64 eln = elen[i];
      if (eln <= 0) continue;
      i64 nvi = -nv[i];
      i64 wnvi = mark - nvi;
      for (i64 q = Cp[i]; q <= Cp[i] + eln - 1; q++) {
        i64 e = Ci[q];
        if (w[e] >= mark) w[e] -= nvi;
        else if (w[e] != 0) w[e] = degree[e] + wnvi;
      }
    }

    /* ---- scan 2: approximate degrees, absorption, hash for supervariables ---- */
    for (i64 pk = pk1; pk < pk2; pk++) {
      i64 i = Ci[pk];
      i64 p1 = Cp[i];
      i64 p2 = p1 + elen[i] - 1;
      i64 pn = p1;
      i64 h = 0, d = 0;
      for (i64 q = p1; q <= p2; q++) {
        i64 e = Ci[q];
        if (w[e] != 0) {
          i64 dext = w[e] - mark;
          if (dext > 0) { d += dext; Ci[pn++] = e; h += e; }
          else { Cp[e] = FLIP(k); w[e] = 0; }   /* aggressive absorption */
        }
      }
      elen[i] = pn - p1 + 1;
      i64 p3 = pn;
      i64 p4 = p1 + len[i];
      for (i64 q = p2 + 1; q < p4; q++) {
        i64 j = Ci[q];
        i64 nvj = nv[j];
        if (nvj <= 0) continue;
        d += nvj;
        Ci[pn++] = j;
        h += j;
      }
      if (d == 0) {             /* mass elimination: i is indistinguishable from k */
        Cp[i] = FLIP(k);
        i64 nvi = -nv[i];
        dk -= nvi; nvk += nvi; nel += nvi;
        nv[i] = 0; elen[i] = -1;
      } else {
        if (d < degree[i]) degree[i] = d;
        Ci[pn] = Ci[p3];
        Ci[p3] = Ci[p1];
        Ci[p1] = k;
        len[i] = pn - p1 + 1;
        h = ((h < 0) ? (-h) : h) % n;
        next[i] = hhead[h];
        hhead[h] = i;
        last[i] = h;
      }
    }
    degree[k] = dk;
    if (dk > lemax) lemax = dk;
    mark = amd_clear_marks(mark + lemax, lemax, w, n);

    /* ---- supervariable detection among the members of Lk ---- */
    for (i64 pk = pk1; pk < pk2; pk++) {
      i64 i = Ci[pk];
      if (nv[i] >= 0) continue;
      i64 h = last[i];
      i = hhead[h];
      hhead[h] = -1;
      for (; i != -1 && next[i] != -1; i = next[i], mark++) {
        i64 ln = len[i], eln = elen[i];
        for (i64 q = Cp[i] + 1; q <= Cp[i] + ln - 1; q++) w[Ci[q]] = mark;
        i64 jlast = i;
        for (i64 j = next[i]; j != -1;) {
          int same = (len[j] == ln) && (elen[j] == eln);
          for (i64 q = Cp[j] + 1; same && q <= Cp[j] + ln - 1; q++)
            if (w[Ci[q]] != mark) same = 0;
          if (same) {
            Cp[j] = FLIP(i);
            nv[i] += nv[j];
            nv[j] = 0;
            elen[j] = -1;
            j = next[j];
            next[jlast] = j;
          } else {
            jlast = j;
            j = next[j];
          }
        }
      }
    }

    /* ---- finalise the new element; put its variables back in the degree lists ---- */
    i64 pf = pk1;
    for (i64 pk = pk1; pk < pk2; pk++) {
      i64 i = Ci[pk];
      i64 nvi = -nv[i];
      if (nvi <= 0) continue;
      nv[i] = nvi;
      i64 d = degree[i] + dk - nvi;
      if (d > n - nel - nvi) d = n - nel - nvi;
      if (head[d] != -1) last[head[d]] = i;
      next[i] = head[d];
      last[i] = -1;
      head[d] = i;
      if (d < mindeg) mindeg = d;
      degree[i] = d;
      Ci[pf++] = i;
    }
    nv[k] = nvk;
    if ((len[k] = pf - pk1) == 0) { Cp[k] = -1; w[k] = 0; }
    if (elenk != 0) cnz = pf;
  }

  /* ---- postorder the assembly tree ---- */
  for (i64 i = 0; i < n; i++) Cp[i] = FLIP(Cp[i]);
  for (i64 j = 0; j <= n; j++) head[j] = -1;
  for (i64 j = n; j >= 0; j--) {           /* absorbed variables under their representative */
    if (nv[j] > 0) continue;
    next[j] = head[Cp[j]];
    head[Cp[j]] = j;
  }
  for (i64 e = n; e >= 0; e--) {           /* elements under their parent element */
    if (nv[e] <= 0) continue;
    if (Cp[e] != -1) { next[e] = head[Cp[e]]; head[Cp[e]] = e; }
  }
  i64 *post = (i64 *)malloc(((size_t)n + 1) * sizeof(i64));
  if (!post) { free(Cp); free(Ci); free(W); free(lastbuf); return -1; }
  i64 kk = 0;
  for (i64 i = 0; i <= n; i++)
    if (Cp[i] == -1) kk = amd_tree_dfs(i, kk, head, next, post, w);
  i64 out = 0;
  for (i64 i = 0; i <= n && out < n; i++)
    if (post[i] != n) perm[out++] = post[i];
  free(post); free(Cp); free(Ci); free(W); free(lastbuf);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* LDLFactStruct restatement                                                             */
/* ------------------------------------------------------------------------------------ */

typedef struct {
  i64 N, nnz;
  i64 *rows, *cols;       /* 0-based copies of the COO lower triangle (rows >= cols)   */
  /* A = triu(sparse(cols, rows, vals)) : CSC, column j holds rows i <= j               */
  i64 nnzA, *Ap, *Ai;
  double *Ax;
  i64 *slot;              /* COO entry -> position in Ax (precomputed variant)          */
  /* ldl_analyze state */
  i64 *P, *pinv;
  i64 *Cp, *Ci, *Cmap;    /* C = triu(P A P'): pattern, and A-nz -> C-nz map            */
  double *Cx;
  i64 *parent, *Lnz, *Lp, *Li;
  double *Lx, *D, *Y;
  i64 *pattern, *flag;
  i64 nnzL;
  double flops;
  int factorized;
} orc_t;

static int cmp_i64(const void *a, const void *b) {
  i64 x = *(const i64 *)a, y = *(const i64 *)b;
  return (x > y) - (x < y);
}

void orc_free(orc_t *h) {
  if (!h) return;
  free(h->rows); free(h->cols); free(h->Ap); free(h->Ai); free(h->Ax); free(h->slot);
  free(h->P); free(h->pinv); free(h->Cp); free(h->Ci); free(h->Cmap); free(h->Cx);
  free(h->parent); free(h->Lnz); free(h->Lp); free(h->Li); free(h->Lx); free(h->D);
  free(h->Y); free(h->pattern); free(h->flag);
  free(h);
}

/*
 * ordering: 0 = AMD (what ldl_analyze does by default), 1 = natural, 2 = user permutation
 * (user_perm[k] = 0-based original index of pivot k).
 * rows1/cols1 are 1-based as handed over by src/CaNNOLeS.jl:276-315.
 * Returns NULL on malformed input (index out of range or strictly-upper triplet).
 */
orc_t *orc_analyze(i64 N, i64 nnz, const i64 *rows1, const i64 *cols1, int ordering,
                   const i64 *user_perm) {
  orc_t *h = (orc_t *)calloc(1, sizeof(orc_t));
  if (!h) return NULL;
  h->N = N; h->nnz = nnz;
  h->rows = (i64 *)malloc((size_t)(nnz + 1) * sizeof(i64));
  h->cols = (i64 *)malloc((size_t)(nnz + 1) * sizeof(i64));
  for (i64 t = 0; t < nnz; t++) {
    i64 r = rows1[t] - 1, c = cols1[t] - 1;
    if (r < 0 || r >= N || c < 0 || c >= N || r < c) { orc_free(h); return NULL; }
    h->rows[t] = r; h->cols[t] = c;
  }
  /* --- sparse(cols, rows, vals, N, N): column = rows[t], row = cols[t]; merge duplicates --- */
  h->Ap = (i64 *)calloc((size_t)N + 1, sizeof(i64));
  i64 *tmpi = (i64 *)malloc((size_t)(nnz + 1) * sizeof(i64));
  i64 *cur = (i64 *)malloc(((size_t)N + 1) * sizeof(i64));
  for (i64 t = 0; t < nnz; t++) h->Ap[h->rows[t] + 1]++;
  for (i64 j = 0; j < N; j++) h->Ap[j + 1] += h->Ap[j];
  memcpy(cur, h->Ap, ((size_t)N + 1) * sizeof(i64));
  for (i64 t = 0; t < nnz; t++) tmpi[cur[h->rows[t]]++] = h->cols[t];
  /* sort each column, unique */
  i64 *newp = (i64 *)malloc(((size_t)N + 1) * sizeof(i64));
  i64 q = 0;
  for (i64 j = 0; j < N; j++) {
    i64 a = h->Ap[j], b = h->Ap[j + 1];
    qsort(tmpi + a, (size_t)(b - a), sizeof(i64), cmp_i64);
    newp[j] = q;
    for (i64 p = a; p < b; p++)
      if (p == a || tmpi[p] != tmpi[p - 1]) tmpi[q++] = tmpi[p];
  }
  newp[N] = q;
  free(h->Ap); h->Ap = newp;
  h->nnzA = q;
  h->Ai = (i64 *)malloc((size_t)(q + 1) * sizeof(i64));
  memcpy(h->Ai, tmpi, (size_t)q * sizeof(i64));
  h->Ax = (double *)calloc((size_t)q + 1, sizeof(double));
  free(tmpi);
  /* slot map by binary search (the same search set_vals! performs per triplet) */
  h->slot = (i64 *)malloc((size_t)(nnz + 1) * sizeof(i64));
  for (i64 t = 0; t < nnz; t++) {
    i64 j = h->rows[t], i = h->cols[t];
    i64 lo = h->Ap[j], hi = h->Ap[j + 1] - 1;
    while (lo < hi) { i64 mid = (lo + hi) >> 1; if (h->Ai[mid] < i) lo = mid + 1; else hi = mid; }
    h->slot[t] = lo;
  }
  /* --- ordering --- */
  h->P = (i64 *)malloc(((size_t)N + 1) * sizeof(i64));
  h->pinv = (i64 *)malloc(((size_t)N + 1) * sizeof(i64));
  if (ordering == 2 && user_perm) {
    memcpy(h->P, user_perm, (size_t)N * sizeof(i64));
  } else if (ordering == 1) {
    for (i64 i = 0; i < N; i++) h->P[i] = i;
  } else {
    /* full symmetric pattern for AMD */
    i64 *Sp = (i64 *)calloc((size_t)N + 2, sizeof(i64));
    for (i64 j = 0; j < N; j++)
      for (i64 p = h->Ap[j]; p < h->Ap[j + 1]; p++) {
        i64 i = h->Ai[p];
        if (i != j) { Sp[i + 1]++; Sp[j + 1]++; }
      }
    for (i64 j = 0; j < N; j++) Sp[j + 1] += Sp[j];
    i64 *Si = (i64 *)malloc((size_t)(Sp[N] + 1) * sizeof(i64));
    memcpy(cur, Sp, ((size_t)N + 1) * sizeof(i64));
    for (i64 j = 0; j < N; j++)
      for (i64 p = h->Ap[j]; p < h->Ap[j + 1]; p++) {
        i64 i = h->Ai[p];
        if (i != j) { Si[cur[i]++] = j; Si[cur[j]++] = i; }
      }
    int rc = orc_amd(N, Sp, Si, h->P);
    free(Sp); free(Si);
    if (rc) { free(cur); orc_free(h); return NULL; }
  }
  for (i64 k = 0; k < N; k++) h->pinv[h->P[k]] = k;
  /* --- C = triu(P A P') by a counting pass over A's columns --- */
  h->Cp = (i64 *)calloc((size_t)N + 1, sizeof(i64));
  h->Ci = (i64 *)malloc((size_t)(h->nnzA + 1) * sizeof(i64));
  h->Cmap = (i64 *)malloc((size_t)(h->nnzA + 1) * sizeof(i64));
  h->Cx = (double *)calloc((size_t)h->nnzA + 1, sizeof(double));
  for (i64 j = 0; j < N; j++)
    for (i64 p = h->Ap[j]; p < h->Ap[j + 1]; p++) {
      i64 i2 = h->pinv[h->Ai[p]], j2 = h->pinv[j];
      h->Cp[(i2 > j2 ? i2 : j2) + 1]++;
    }
  for (i64 j = 0; j < N; j++) h->Cp[j + 1] += h->Cp[j];
  memcpy(cur, h->Cp, ((size_t)N + 1) * sizeof(i64));
  for (i64 j = 0; j < N; j++)
    for (i64 p = h->Ap[j]; p < h->Ap[j + 1]; p++) {
      i64 i2 = h->pinv[h->Ai[p]], j2 = h->pinv[j];
      i64 c = i2 > j2 ? i2 : j2, r = i2 > j2 ? j2 : i2;
      i64 dst = cur[c]++;
      h->Ci[dst] = r;
      h->Cmap[p] = dst;
    }
  free(cur);
  /* --- ldl_symbolic: elimination tree and column counts of L --- */
  h->parent = (i64 *)malloc(((size_t)N + 1) * sizeof(i64));
  h->Lnz = (i64 *)calloc((size_t)N + 1, sizeof(i64));
  h->flag = (i64 *)malloc(((size_t)N + 1) * sizeof(i64));
  h->Lp = (i64 *)malloc(((size_t)N + 1) * sizeof(i64));
  for (i64 k = 0; k < N; k++) {
    h->parent[k] = -1;
    h->flag[k] = k;
    for (i64 p = h->Cp[k]; p < h->Cp[k + 1]; p++) {
      i64 i = h->Ci[p];
      for (; i < k && h->flag[i] != k; i = h->parent[i]) {
        if (h->parent[i] == -1) h->parent[i] = k;
        h->Lnz[i]++;
        h->flag[i] = k;
      }
    }
  }
  h->Lp[0] = 0;
  double fl = 0;
  for (i64 k = 0; k < N; k++) {
    h->Lp[k + 1] = h->Lp[k] + h->Lnz[k];
    double c = (double)h->Lnz[k];
    fl += c * c + 3.0 * c;
  }
  h->nnzL = h->Lp[N];
  h->flops = fl;
  h->Li = (i64 *)malloc((size_t)(h->nnzL + 1) * sizeof(i64));
  h->Lx = (double *)malloc((size_t)(h->nnzL + 1) * sizeof(double));
  h->D = (double *)calloc((size_t)N + 1, sizeof(double));
  h->Y = (double *)calloc((size_t)N + 1, sizeof(double));
  h->pattern = (i64 *)malloc(((size_t)N + 1) * sizeof(i64));
  return h;
}

/* set_vals!, faithful variant: zero, then for every triplet a binary search in column
 * rows[t] and += (src/solver_types.jl:53-59). */
void orc_set_vals_search(orc_t *h, const double *vals) {
  memset(h->Ax, 0, (size_t)h->nnzA * sizeof(double));
  for (i64 t = 0; t < h->nnz; t++) {
    i64 j = h->rows[t], i = h->cols[t];
    i64 lo = h->Ap[j], hi = h->Ap[j + 1] - 1;
    while (lo < hi) { i64 mid = (lo + hi) >> 1; if (h->Ai[mid] < i) lo = mid + 1; else hi = mid; }
    h->Ax[lo] += vals[t];
  }
}

/* set_vals!, "fair" variant: same sums in the same order through the precomputed slot map */
void orc_set_vals(orc_t *h, const double *vals) {
  memset(h->Ax, 0, (size_t)h->nnzA * sizeof(double));
  for (i64 t = 0; t < h->nnz; t++) h->Ax[h->slot[t]] += vals[t];
}

/* ldl_factorize!: numeric up-looking LDL^T of C = triu(P A P').  Returns N on success or the
 * 0-based index of the first exact-zero pivot (entries of D beyond it keep stale values, as in
 * the reference, whose caller inspects d regardless). */
i64 orc_ldl_factorize(orc_t *h) {
  i64 N = h->N;
  i64 *Lp = h->Lp, *Li = h->Li, *parent = h->parent, *Lnz = h->Lnz, *flag = h->flag,
      *pattern = h->pattern;
  double *Lx = h->Lx, *D = h->D, *Y = h->Y;
  for (i64 p = 0; p < h->nnzA; p++) h->Cx[h->Cmap[p]] = h->Ax[p];
  h->factorized = 0;
  memset(Lnz, 0, (size_t)N * sizeof(i64)); /* columns past a breakdown stay empty */
  for (i64 k = 0; k < N; k++) {
    Y[k] = 0.0;
    i64 top = N;
    flag[k] = k;
    Lnz[k] = 0;
    for (i64 p = h->Cp[k]; p < h->Cp[k + 1]; p++) {
      i64 i = h->Ci[p];
      Y[i] += h->Cx[p];
      i64 len = 0;
      for (; flag[i] != k; i = parent[i]) {
        pattern[len++] = i;
        flag[i] = k;
      }
      while (len > 0) pattern[--top] = pattern[--len];
    }
    D[k] = Y[k];
    Y[k] = 0.0;
    for (; top < N; top++) {
      i64 i = pattern[top];
      double yi = Y[i];
      Y[i] = 0.0;
      i64 p2 = Lp[i] + Lnz[i];
      i64 p;
      for (p = Lp[i]; p < p2; p++) Y[Li[p]] -= Lx[p] * yi;
      double l_ki = yi / D[i];
      D[k] -= l_ki * yi;
      Li[p] = k;
      Lx[p] = l_ki;
      Lnz[i]++;
    }
    if (D[k] == 0.0) return k;
  }
  h->factorized = 1;
  return N;
}

/* inertia loop of try_to_factorize (src/solver_types.jl:90-96) + zero/neg split */
void orc_inertia(const orc_t *h, double eig_tol, i64 *npos, i64 *nzero, i64 *nneg) {
  i64 pos = 0, zer = 0, neg = 0;
  for (i64 i = 0; i < h->N; i++) {
    double di = h->D[i];
    pos += di > eig_tol;
    zer += fabs(di) <= eig_tol;
    neg += di < -eig_tol;
  }
  *npos = pos; *nzero = zer; *nneg = neg;
}

/* try_to_factorize (src/solver_types.jl:79-98). use_search=1 times the reference's real
 * set_vals! behaviour. */
int orc_try_to_factorize(orc_t *h, const double *vals, i64 nvar, i64 nequ, i64 ncon,
                         double eig_tol, int use_search) {
  (void)nequ; (void)ncon;
  if (use_search) orc_set_vals_search(h, vals); else orc_set_vals(h, vals);
  orc_ldl_factorize(h);
  i64 pos, zer, neg;
  orc_inertia(h, eig_tol, &pos, &zer, &neg);
  return pos == nvar && zer == 0;
}

/* solve_ldl! (src/solver_types.jl:69-77): d = -(K^{-1} rhs) */
int orc_solve_ldl(orc_t *h, const double *rhs, double *d) {
  i64 N = h->N;
  double *y = h->Y;
  for (i64 k = 0; k < N; k++) y[k] = rhs[h->P[k]];
  for (i64 j = 0; j < N; j++) {
    double yj = y[j];
    i64 p2 = h->Lp[j] + h->Lnz[j];
    for (i64 p = h->Lp[j]; p < p2; p++) y[h->Li[p]] -= h->Lx[p] * yj;
  }
  for (i64 j = 0; j < N; j++) y[j] /= h->D[j];
  for (i64 j = N - 1; j >= 0; j--) {
    double yj = y[j];
    i64 p2 = h->Lp[j] + h->Lnz[j];
    for (i64 p = h->Lp[j]; p < p2; p++) yj -= h->Lx[p] * y[h->Li[p]];
    y[j] = yj;
  }
  for (i64 k = 0; k < N; k++) d[h->P[k]] = y[k];
  for (i64 k = 0; k < N; k++) d[k] = -d[k];
  for (i64 k = 0; k < N; k++) y[k] = 0.0;
  return 1;
}

/* y = K x using the assembled upper CSC (for residual checks in tests) */
void orc_matvec(const orc_t *h, const double *x, double *y) {
  for (i64 i = 0; i < h->N; i++) y[i] = 0.0;
  for (i64 j = 0; j < h->N; j++)
    for (i64 p = h->Ap[j]; p < h->Ap[j + 1]; p++) {
      i64 i = h->Ai[p];
      y[i] += h->Ax[p] * x[j];
      if (i != j) y[j] += h->Ax[p] * x[i];
    }
}

/* accessors */
i64 orc_N(const orc_t *h) { return h->N; }
i64 orc_nnzA(const orc_t *h) { return h->nnzA; }
i64 orc_nnzL(const orc_t *h) { return h->nnzL; }
double orc_flops(const orc_t *h) { return h->flops; }
const i64 *orc_Ap(const orc_t *h) { return h->Ap; }
const i64 *orc_Ai(const orc_t *h) { return h->Ai; }
const double *orc_Ax(const orc_t *h) { return h->Ax; }
const i64 *orc_slot(const orc_t *h) { return h->slot; }
const i64 *orc_perm(const orc_t *h) { return h->P; }
const double *orc_D(const orc_t *h) { return h->D; }
const i64 *orc_parent(const orc_t *h) { return h->parent; }
const i64 *orc_Lp(const orc_t *h) { return h->Lp; }

double orc_wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Timed loop for the CPU baseline: `reps` x (set_vals! + factorize + inertia + solve).
 * Returns seconds per repetition; phase split in t3[0..2] = (assemble, factorize, solve). */
double orc_time_factor_solve(orc_t *h, const double *vals, const double *rhs, double *d,
                             i64 nvar, double eig_tol, int reps, int use_search, double *t3,
                             int *ok_out) {
  double ta = 0, tf = 0, ts = 0;
  int ok = 1;
  for (int r = 0; r < reps; r++) {
    double t0 = orc_wtime();
    if (use_search) orc_set_vals_search(h, vals); else orc_set_vals(h, vals);
    double t1 = orc_wtime();
    orc_ldl_factorize(h);
    i64 pos, zer, neg;
    orc_inertia(h, eig_tol, &pos, &zer, &neg);
    ok = ok && (pos == nvar && zer == 0);
    double t2 = orc_wtime();
    orc_solve_ldl(h, rhs, d);
    double t3e = orc_wtime();
    ta += t1 - t0; tf += t2 - t1; ts += t3e - t2;
  }
  if (t3) { t3[0] = ta / reps; t3[1] = tf / reps; t3[2] = ts / reps; }
  if (ok_out) *ok_out = ok;
  return (ta + tf + ts) / reps;
}
