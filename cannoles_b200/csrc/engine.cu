// engine.cu -- host orchestration of the single-system engine (see engine.h, kernels.cuh).
#include "engine.h"
#include "kernels.cuh"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstring>

namespace b2 {

thread_local char g_last_error[512] = "";

namespace {

template <typename T>
int upload(T** dptr, const std::vector<T>& v, double& bytes, size_t min_elems = 1) {
  size_t n = std::max(v.size(), min_elems);
  B2_CUDA_OK(cudaMalloc((void**)dptr, n * sizeof(T)));
  if (!v.empty()) B2_CUDA_OK(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  bytes += (double)(n * sizeof(T));
  return 0;
}

template <typename T>
int dalloc(T** dptr, size_t n, double& bytes) {
  n = std::max<size_t>(n, 1);
  B2_CUDA_OK(cudaMalloc((void**)dptr, n * sizeof(T)));
  bytes += (double)(n * sizeof(T));
  return 0;
}

int g_small_c0_max = 16;   // fronts up to this order: one warp per front, four fronts per CTA (developer knob B2_SMALL_C0_MAX)
int small_class(int m) { return m <= g_small_c0_max ? 0 : (m <= 40 ? 1 : (m <= 72 ? 2 : 3)); }
int solve_class(int m) { return m <= 32 ? 0 : (m <= 128 ? 1 : (m <= 512 ? 2 : 3)); }

double wall() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

int Engine::build_plan() {
  if (const char* e = getenv("B2_SMALL_C0_MAX")) g_small_c0_max = std::max(8, std::min(40, atoi(e)));
  const Symbolic& S = sym;
  if (S.uptr[S.nsuper] >= (int64_t)INT32_MAX) {
    snprintf(g_last_error, sizeof(g_last_error), "update-vector storage exceeds int32 offsets");
    return -1;
  }
  std::vector<int32_t> items;
  std::vector<int64_t> asm_cptr(1, 0);   // destination column of a tiled front -> (child, child column) pairs
  std::vector<int32_t> asm_ent, asm_rc;
  std::vector<int64_t> asm_off, sb_ptr(1, 0);
  std::vector<int32_t> sb_src, sb_flag(S.nsuper, 0);   // big-front solve: CSR gather of the child updates, flag offsets
  int64_t n_gather_chunks = 0;
  nsflag = 0;
  n_stk = 0;
  fact_launches.clear(); fwd_launches.clear(); bwd_launches.clear();
  n_small = n_large = 0;
  std::vector<char> front_dag(S.nsuper, 0);
  auto front_m = [&](int s) { return (int)(S.rptr[s + 1] - S.rptr[s]); };
  auto front_w = [&](int s) { return (int)(S.scol[s + 1] - S.scol[s]); };
  // Cross-level dataflow: the top of the assembly tree -- the highest run of levels with at most
  // dag_level_max fronts each, if it holds a front of the tiled path at all -- is ONE k_front_dag
  // launch: every front of those levels is cut into NB x NB tile tasks, the extend-add is a task
  // kind, and a parent waits for ITS children only (counters), not for a level.
  dag_from_level = S.nlevels;
  if (use_dag) {
    int l = S.nlevels;
    while (l > 0 && S.level_ptr[l] - S.level_ptr[l - 1] <= dag_level_max) l--;
    bool any_large = false;
    for (int q = S.level_ptr[l]; q < S.level_ptr[S.nlevels] && !any_large; q++)
      any_large = front_m(S.level_sn[q]) > (int)small_max_m;
    if (any_large) dag_from_level = l;
  }
  struct DagFront { int s, m, w, np, nrb; int32_t fb, df, tbase; };
  struct TlEnt { int64_t gid; int32_t ce, ia, iz, ja, jz, rlo, clo; };
  struct FlEnt { int64_t gid; int32_t dest, seq, isA, src; };
  struct DTask { double key; int32_t rec[8]; };
  std::vector<DTask> dtasks;                        // tasks of the dataflow launch with their schedule keys
  std::vector<double> f_done, f_kdone;              // per front of the launch: modelled completion time / key of its last tile
  std::vector<FlEnt> fl_tmp;                        // flat extend-add elements of the front in hand
  std::vector<int32_t> fl_cnt, fl_ent;
  const bool flat_ok = S.cb_store < (int64_t)INT32_MAX && S.nnzA < (int64_t)INT32_MAX;
  std::vector<TlEnt> tl_tmp;                        // (tile, child) pairs in front / child order
  std::vector<int32_t> tl_cnt;
  std::vector<DagFront> dagf;                       // fronts of the dataflow launch, in level order
  std::vector<int32_t> dag_of(S.nsuper, -1);        // front -> its record
  std::vector<int32_t> items_dag;                   // task records (8 ints each), level by level
  std::vector<int32_t> dfr;
  Launch G; G.kind = LK_DAG;
  std::vector<size_t> fstart, sstart, bstart;   // first launch of every level in the three lists
  for (int l = 0; l < S.nlevels; l++) {
    fstart.push_back(fact_launches.size()); sstart.push_back(fwd_launches.size()); bstart.push_back(bwd_launches.size());
    std::vector<int32_t> small[4], large, sol[4], big, tiny[2], tsol[2];
    int small_mmax[4] = {0, 0, 0, 0}, sol_mmax[4] = {0, 0, 0, 0};
    const bool dagl = l >= dag_from_level;          // a level of the dataflow launch: no other factorization launch
    for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; q++) {
      int s = S.level_sn[q];
      int m = front_m(s);
      if (m <= tiny_max_m && m <= (int)small_max_m) {   // one thread per front, factorization and solves
        tiny[m <= 4 ? 0 : 1].push_back(s);
        if (dagl) { large.push_back(s); n_large++; } else n_small++;
        continue;
      }
      if (dagl) {
        large.push_back(s);
        n_large++;
      } else if (m <= (int)small_max_m) {
        int c = small_class(m);
        small[c].push_back(s);
        small_mmax[c] = std::max(small_mmax[c], m);
        n_small++;
      } else {
        large.push_back(s);
        n_large++;
      }
      if (m > (int)solve_big_m) {
        big.push_back(s);
      } else if (m <= tiny_solve_max_m) {
        tsol[m <= 16 ? 0 : 1].push_back(s);
      } else {
        int c = solve_class(m);
        sol[c].push_back(s);
        sol_mmax[c] = std::max(sol_mmax[c], m);
      }
    }
    for (int c = 0; c < 2; c++) {
      if (tiny[c].empty()) continue;
      Launch L; L.kind = LK_FRONT_TINY; L.cls = c; L.off = (int64_t)items.size(); L.count = (int)tiny[c].size();
      items.insert(items.end(), tiny[c].begin(), tiny[c].end());
      if (!dagl) fact_launches.push_back(L);
      L.kind = LK_FWD_TINY; fwd_launches.push_back(L);
      L.kind = LK_BWD_TINY; bwd_launches.push_back(L);
    }
    for (int c = 0; c < 2; c++) {
      if (tsol[c].empty()) continue;
      Launch L; L.kind = LK_FWD_TINY; L.cls = 2 + c; L.off = (int64_t)items.size(); L.count = (int)tsol[c].size();
      items.insert(items.end(), tsol[c].begin(), tsol[c].end());
      fwd_launches.push_back(L);
      L.kind = LK_BWD_TINY; bwd_launches.push_back(L);
    }
    for (int c = 0; c < 4; c++) {
      if (!small[c].empty()) {
        Launch L; L.kind = LK_FRONT_SMALL; L.cls = c; L.off = (int64_t)items.size();
        L.count = (int)small[c].size();
        L.smem = (int)(((size_t)small_mmax[c] * small_mmax[c] + 2 * (size_t)small_mmax[c]) * sizeof(double));
        items.insert(items.end(), small[c].begin(), small[c].end());
        fact_launches.push_back(L);
      }
      if (!sol[c].empty()) {
        Launch L; L.kind = LK_FWD; L.cls = c; L.off = (int64_t)items.size();
        L.count = (int)sol[c].size();
        L.smem = (int)(((size_t)sol_mmax[c] + SNB) * sizeof(double));
        items.insert(items.end(), sol[c].begin(), sol[c].end());
        fwd_launches.push_back(L);
        L.kind = LK_BWD;
        bwd_launches.push_back(L);
      }
    }
    if (!big.empty()) {   // multi-CTA solves: flag-chained chunks, items ordered so that waits look back
      Launch F; F.kind = LK_FWD_BIG; F.off = (int64_t)items.size(); F.jb = n_stk++;
      int rmax = 0;
      for (int s : big) {
        int w = front_w(s), m = front_m(s);
        int nblk = (w + SB - 1) / SB, nbel = (m - w + SB - 1) / SB;
        sb_flag[s] = (int32_t)nsflag;
        nsflag += nblk;
        rmax = std::max(rmax, m - w);
        // per (chunk, row of the chunk): the child update entries that land there, in child order
        std::vector<std::vector<int32_t>> bucket((size_t)(nblk + nbel) * SB);
        for (int q = S.child_ptr[s]; q < S.child_ptr[s + 1]; q++) {
          int c = S.child_idx[q];
          int wc = front_w(c);
          const int32_t* relc = &S.rel[S.rptr[c] + wc];
          int rc = front_m(c) - wc;
          for (int k = 0; k < rc; k++) {
            int row = relc[k];
            int ch = row < w ? row / SB : nblk + (row - w) / SB;
            int rr = row < w ? row % SB : (row - w) % SB;
            bucket[(size_t)ch * SB + rr].push_back((int32_t)(S.uptr[c] + k));
          }
        }
        for (int c = 0; c < nblk + nbel; c++) {
          for (int rr = 0; rr < SB; rr++) {
            const auto& bk = bucket[(size_t)c * SB + rr];
            sb_src.insert(sb_src.end(), bk.begin(), bk.end());
            sb_ptr.push_back((int64_t)sb_src.size());
          }
          sb_ptr.push_back((int64_t)sb_src.size());   // pad to SB + 1 pointers per chunk
          // (pointer layout per chunk g: sb_ptr[g*(SB+1) + r] .. [+ r + 1]; the leading 0 belongs to chunk 0)
          items.push_back(s); items.push_back(c); items.push_back((int)n_gather_chunks); items.push_back(0);
          n_gather_chunks++;
          F.count++;
        }
      }
      fwd_launches.push_back(F);
      Launch Bk; Bk.kind = LK_BWD_BIG; Bk.off = (int64_t)items.size(); Bk.jb = n_stk++;
      Bk.smem = (int)((size_t)std::max(rmax, 1) * sizeof(double));
      for (int s : big) {
        int nblk = (front_w(s) + SB - 1) / SB;
        for (int c = nblk - 1; c >= 0; c--) { items.push_back(s); items.push_back(c); Bk.count++; }
      }
      bwd_launches.push_back(Bk);
    }
    if (large.empty()) continue;
    // extend-add items of one tiled front (8 ints each, see assemble_tile_body); for a front of the
    // dataflow launch also the expected number of items per 64-column group of the front
    auto asm_items = [&](int s, std::vector<int32_t>& out, int df, int32_t* expect) -> int {
      int cnt = 0;
      int m = front_m(s);
      const int acols = m <= 128 ? 16 : 8, arows = ASM_TILE / acols;   // tile shape of this front
      int ncb = (m + acols - 1) / acols, nrb = (m + arows - 1) / arows;
      // per destination column: the (child descriptor, child column) pairs that land on it, children
      // ascending (the extend-add order)
      const int64_t gbase = (int64_t)asm_cptr.size() - 1;   // global id of this front's column 0
      std::vector<std::vector<int32_t>> percol(m);
      for (int q = S.child_ptr[s]; q < S.child_ptr[s + 1]; q++) {
        int c = S.child_idx[q];
        int wc = front_w(c);
        const int32_t* relc = &S.rel[S.rptr[c] + wc];
        int rc = front_m(c) - wc;
        if (rc == 0) continue;
        const int32_t ce = (int32_t)asm_rc.size();
        asm_rc.push_back(rc);
        asm_off.push_back(S.rptr[c] + wc); asm_off.push_back(S.cbptr[c]);
        for (int j = 0; j < rc; j++) { percol[relc[j]].push_back(ce); percol[relc[j]].push_back(j); }
      }
      for (int J = 0; J < m; J++) {
        asm_ent.insert(asm_ent.end(), percol[J].begin(), percol[J].end());
        asm_cptr.push_back((int64_t)asm_ent.size() / 2);
      }
      const int32_t* apos = S.amap_pos.data() + S.amap_ptr[s];
      const int64_t na = S.amap_ptr[s + 1] - S.amap_ptr[s];
      const int w = front_w(s);
      for (int cb = 0; cb < ncb; cb++) {
        // A entries of pivot columns [j0, min(je, w)): positions row + col * m, sorted
        int j0 = cb * acols, je = std::min(j0 + acols, m);
        int qa = 0, qb = 0;
        if (j0 < w) {
          qa = (int)(std::lower_bound(apos, apos + na, j0 * m) - apos);
          qb = (int)(std::lower_bound(apos, apos + na, std::min(je, w) * m) - apos);
        }
        for (int rb = 0; rb < nrb; rb++) {
          if ((rb + 1) * arows <= cb * acols) continue;   // tile entirely above the diagonal
          out.push_back(s); out.push_back(j0); out.push_back((rb * arows) | (3 << 16)); out.push_back((int32_t)(gbase + j0));
          out.push_back(qa); out.push_back(qb); out.push_back(acols | (arows << 16)); out.push_back(df);
          if (expect) expect[j0 / 64]++;
          cnt++;
        }
      }
      return cnt;
    };
    if (dagl) {
      // tasks over the NB x NB tiles (I, J), I >= J, of every front of the level (see k_front_dag)
      std::vector<DagFront> fg;
      for (int s : large) {
        int w = front_w(s), m = front_m(s);
        int np = (w + NB - 1) / NB, nrb = np + (m - w + NB - 1) / NB;
        if (m >= 65536 || ntflag + (int64_t)np * (nrb + 1) >= (int64_t)INT32_MAX) { snprintf(g_last_error, sizeof(g_last_error), "too many tiles"); return -1; }
        fg.push_back({s, m, w, np, nrb, (int32_t)ntflag, 0, 0});
        ntflag += (int64_t)np * (nrb + 1);   // one flag per tile of the pivot columns + one per ypre task
      }
      std::stable_sort(fg.begin(), fg.end(), [](const DagFront& a, const DagFront& b) { return a.m > b.m; });
      for (DagFront& f : fg) {
        f.df = (int32_t)dagf.size();
        dag_of[f.s] = f.df;
        front_dag[f.s] = 1;
        int nch = 0;
        for (int q = S.child_ptr[f.s]; q < S.child_ptr[f.s + 1]; q++) nch += dag_of[S.child_idx[q]] >= 0;
        const int nq = f.nrb - f.np;
        dfr.push_back(0); dfr.push_back(nq * (nq + 1) / 2); dfr.push_back(-1); dfr.push_back(nch);
        // extend-add lists: per tile (I, J) of the front the children that land on it, with the child
        // rows / columns that fall into row block I / column block J (rel[] is ascending: ranges)
        f.tbase = (int32_t)tl_cnt.size();
        tl_cnt.resize(tl_cnt.size() + (size_t)f.nrb * f.nrb, 0);
        auto blk_lo = [&](int b) { return b < f.np ? b * NB : f.w + (b - f.np) * NB; };
        auto blk_of = [&](int row) { return row < f.w ? row / NB : f.np + (row - f.w) / NB; };
        fl_tmp.clear();
        std::vector<int> ca(f.nrb + 1);
        for (int q = S.child_ptr[f.s]; q < S.child_ptr[f.s + 1]; q++) {
          const int c = S.child_idx[q];
          const int wc = front_w(c), rc = front_m(c) - wc;
          if (rc == 0) continue;
          const int32_t* relc = &S.rel[S.rptr[c] + wc];
          if (rc <= DAG_SMALL_RC && flat_ok) {
            // a small child: its entries go to the flat per-tile lists (sorted by destination below)
            for (int jc = 0; jc < rc; jc++) {
              const int bj = blk_of(relc[jc]);
              for (int ic = jc; ic < rc; ic++) {
                const int bi = blk_of(relc[ic]);
                fl_tmp.push_back({(int64_t)f.tbase + (int64_t)bj * f.nrb + bi,
                                  (relc[jc] - blk_lo(bj)) * DAG_LDT + (relc[ic] - blk_lo(bi)), 1 + q - S.child_ptr[f.s], 0,
                                  (int32_t)(S.cbptr[c] + ic + (int64_t)jc * rc)});
              }
            }
            continue;
          }
          const int32_t ce = (int32_t)asm_rc.size();
          asm_rc.push_back(rc);
          asm_off.push_back(S.rptr[c] + wc); asm_off.push_back(S.cbptr[c]);
          for (int b = 0, k = 0; b <= f.nrb; b++) {       // ca[b] = first child row in block >= b
            const int lo = b < f.nrb ? blk_lo(b) : f.m;
            while (k < rc && relc[k] < lo) k++;
            ca[b] = k;
          }
          for (int bj = 0; bj < f.nrb; bj++) {
            if (ca[bj] == ca[bj + 1]) continue;
            for (int bi = bj; bi < f.nrb; bi++) {
              if (ca[bi] == ca[bi + 1]) continue;
              tl_tmp.push_back({(int64_t)f.tbase + (int64_t)bj * f.nrb + bi, ce, ca[bi], ca[bi + 1], ca[bj], ca[bj + 1], blk_lo(bi), blk_lo(bj)});
              tl_cnt[(size_t)f.tbase + (size_t)bj * f.nrb + bi]++;
            }
          }
        }
        if (getenv("B2_DAG_DEBUG") && f.m > 600) {
          int nchild = S.child_ptr[f.s + 1] - S.child_ptr[f.s], small = 0, t00 = tl_cnt[f.tbase], tmax = 0;
          long long e00 = 0;
          for (int q = S.child_ptr[f.s]; q < S.child_ptr[f.s + 1]; q++) small += front_m(S.child_idx[q]) - front_w(S.child_idx[q]) <= 32;
          for (int b = 0; b < f.nrb * f.nrb; b++) tmax = std::max(tmax, (int)tl_cnt[f.tbase + b]);
          for (size_t t = 0; t < tl_tmp.size(); t++) if (tl_tmp[t].gid == f.tbase) e00 += (long long)(tl_tmp[t].iz - tl_tmp[t].ia) * (tl_tmp[t].jz - tl_tmp[t].ja);
          fprintf(stderr, "dagfront level %d m %d w %d: children %d (rc <= 32: %d), pairs on tile (0,0) %d (%lld entries), max pairs per tile %d\n", l, f.m, f.w, nchild, small, t00, e00, tmax);
        }
        // the A entries of the front (position = row + col * m inside the panel)
        {
          const int32_t* apos = S.amap_pos.data() + S.amap_ptr[f.s];
          const int32_t* aslot = S.amap_slot.data() + S.amap_ptr[f.s];
          const int64_t na = S.amap_ptr[f.s + 1] - S.amap_ptr[f.s];
          if (flat_ok) {
            for (int64_t q = 0; q < na; q++) {
              const int j = apos[q] / f.m, i = apos[q] - j * f.m;
              const int bj = blk_of(j), bi = blk_of(i);
              fl_tmp.push_back({(int64_t)f.tbase + (int64_t)bj * f.nrb + bi, (j - blk_lo(bj)) * DAG_LDT + (i - blk_lo(bi)), 0, 1, aslot[q]});
            }
          } else if (na > 0) { snprintf(g_last_error, sizeof(g_last_error), "contribution-block storage exceeds int32 offsets"); return -1; }
        }
        // sort by (tile, destination, A first, children ascending); the first element of a run carries
        // its length; runs are cut at the staging rounds of the kernel
        std::sort(fl_tmp.begin(), fl_tmp.end(), [](const FlEnt& a, const FlEnt& b) {
          return a.gid != b.gid ? a.gid < b.gid : (a.dest != b.dest ? a.dest < b.dest : a.seq < b.seq); });
        fl_cnt.resize(fl_cnt.size() + (size_t)f.nrb * f.nrb, 0);
        for (size_t a = 0; a < fl_tmp.size();) {
          size_t t1 = a;
          while (t1 < fl_tmp.size() && fl_tmp[t1].gid == fl_tmp[a].gid) t1++;
          fl_cnt[(size_t)fl_tmp[a].gid] = (int32_t)(t1 - a);
          for (size_t e = a; e < t1;) {
            size_t r1 = e + 1;
            while (r1 < t1 && fl_tmp[r1].dest == fl_tmp[e].dest && ((r1 - a) % DAG_STAGE) != 0) r1++;
            for (size_t k = e; k < r1; k++) {
              fl_ent.push_back(fl_tmp[k].dest | (fl_tmp[k].isA << 13) | (k == e ? (int32_t)((r1 - e) << 14) : 0));
              fl_ent.push_back(fl_tmp[k].src);
            }
            e = r1;
          }
          a = t1;
        }
        dagf.push_back(f);
      }
      // ticket order inside a level = waves of the tile DAG over all its fronts (a topological
      // order): wave 2d holds what can start once pivot block d-1 is factored -- the chain task of
      // block d (first: it is the critical path) and the tiles of column d-1; 2d+1 the ypre task of
      // block d+1; the tiles of the contribution block follow the last column (wave 2 np + 1)
      struct TK { int key, pri; int32_t s, I, code, fb, df, tbase; };
      std::vector<TK> tks;
      for (const DagFront& f : fg) {
        const int np = f.np;
        for (int J = 0; J < f.nrb; J++)
          for (int I = J; I < f.nrb; I++) {
            if (J >= np) { tks.push_back({2 * np + 1, 1, f.s, I, J, f.fb, f.df, f.tbase}); continue; }
            if (I == J) {
              if (J == 0) tks.push_back({0, 0, f.s, 0, 0, f.fb, f.df, f.tbase});
              continue;                                      // J > 0: part of the chain task of block J
            }
            if (I == J + 1 && I < np) {                      // chain task: tiles (I, I-1) and (I, I) ...
              tks.push_back({2 * I, 0, f.s, I, I | (1 << 16), f.fb, f.df, f.tbase});
              // ... after the task that applies the pivot blocks p < I-1 to (I, I)
              if (I >= 2) tks.push_back({2 * (I - 1) + 1, 1, f.s, I, I | (2 << 16), f.fb, f.df, f.tbase});
              continue;
            }
            tks.push_back({2 * (J + 1), 1, f.s, I, J, f.fb, f.df, f.tbase});
          }
      }
      std::stable_sort(tks.begin(), tks.end(), [](const TK& a, const TK& b) { return a.key != b.key ? a.key < b.key : a.pri < b.pri; });
      // Ticket order of the whole launch = a simulated schedule.  The tasks are walked in the order
      // above (topological) with a cost model (us) and unlimited workers: `fin` = earliest finish of
      // a task given when its inputs appear, key = fin - own work = the latest start that does not
      // delay it (a left-looking tile started earlier would only hold a CTA while it waits for pivot
      // blocks).  The key is then raised above the keys of everything the task waits for, so that
      // sorting by key stays a topological order: a CTA only ever waits for smaller tickets.
      {
        // (cost model in us; T_UPD is per 64-column pivot block applied to a tile.  Developer knobs
        // B2_DAG_TUPD / B2_DAG_TASM override the two that the task traces disagree with most.)
        double T_ASM = 8, T_UPD = 2.5;
        constexpr double T_LDL = 12.5, T_SUB = 5, T_POST = 2, T_ST = 1, T_EPS = 1e-3;
        if (const char* e = getenv("B2_DAG_TUPD")) T_UPD = atof(e);
        if (const char* e = getenv("B2_DAG_TASM")) T_ASM = atof(e);
        struct FS { double R = 0, KR = 0; std::vector<double> F, KF, Y, KY; };
        const int df0 = fg.empty() ? 0 : fg[0].df;
        std::vector<FS> fs(fg.size());
        f_done.resize(dagf.size(), 0.0); f_kdone.resize(dagf.size(), 0.0);
        for (size_t i = 0; i < fg.size(); i++) {
          const DagFront& f = fg[i];
          fs[i].F.assign((size_t)f.np * f.nrb, 0.0); fs[i].KF.assign((size_t)f.np * f.nrb, 0.0);
          fs[i].Y.assign(f.np + 1, 0.0); fs[i].KY.assign(f.np + 1, 0.0);
          for (int q = S.child_ptr[f.s]; q < S.child_ptr[f.s + 1]; q++) {
            const int dc = dag_of[S.child_idx[q]];
            if (dc >= 0) { fs[i].R = std::max(fs[i].R, f_done[dc]); fs[i].KR = std::max(fs[i].KR, f_kdone[dc]); }
          }
        }
        for (const TK& t : tks) {
          const DagFront& f = fg[t.df - df0];
          FS& X = fs[t.df - df0];
          const int np = f.np, nrb = f.nrb, kind = t.code >> 16, J = t.code & 0xffff, I = t.I;
          double tc = X.R + T_ASM, work = T_ASM, kdep = X.KR;
          auto need = [&](int i, int p) { tc = std::max(tc, X.F[i + (size_t)p * nrb]); kdep = std::max(kdep, X.KF[i + (size_t)p * nrb]); };
          auto upd = [&]() { tc += T_UPD; work += T_UPD; };
          double key = 0;
          if (kind == 0 && J >= np) {                          // tile of the contribution block
            for (int p2 = 0; p2 < np; p2++) { need(I, p2); need(J, p2); upd(); }
            tc += T_ST; work += T_ST;
            key = std::max(tc - work, kdep + T_EPS);
            f_done[t.df] = std::max(f_done[t.df], tc); f_kdone[t.df] = std::max(f_kdone[t.df], key);
          } else if (kind == 0 && I == J) {                    // the first diagonal tile
            tc += T_LDL + T_POST; work += T_LDL + T_POST;
            key = std::max(tc - work, kdep + T_EPS);
            X.F[0] = tc; X.KF[0] = key;
          } else if (kind == 0) {                              // substitution tile (I, J)
            for (int p2 = 0; p2 < J; p2++) { need(I, p2); need(J, p2); upd(); }
            need(J, J); tc += T_SUB; work += T_SUB;
            key = std::max(tc - work, kdep + T_EPS);
            X.F[I + (size_t)J * nrb] = tc; X.KF[I + (size_t)J * nrb] = key;
          } else if (kind == 1) {                              // chain task of block J
            for (int p2 = 0; p2 < J - 1; p2++) { need(J, p2); need(J - 1, p2); upd(); }
            if (J >= 2) { tc = std::max(tc, X.Y[J]); kdep = std::max(kdep, X.KY[J]); } else { tc += T_ASM; work += T_ASM; }
            need(J - 1, J - 1); tc += T_SUB; work += T_SUB;
            const double tsub = tc;
            tc += T_UPD + T_LDL + T_POST; work += T_UPD + T_LDL + T_POST;
            key = std::max(tc - work, kdep + T_EPS);
            X.F[J + (size_t)(J - 1) * nrb] = tsub; X.KF[J + (size_t)(J - 1) * nrb] = key;
            X.F[J + (size_t)J * nrb] = tc; X.KF[J + (size_t)J * nrb] = key;
          } else {                                             // ypre task of block J
            for (int p2 = 0; p2 < J - 1; p2++) { need(J, p2); upd(); }
            tc += T_ST; work += T_ST;
            key = std::max(tc - work, kdep + T_EPS);
            X.Y[J] = tc; X.KY[J] = key;
          }
          DTask d; d.key = key;
          const int32_t rec[8] = {t.s, t.I, t.code, t.fb, t.tbase, 0, 0, t.df};
          memcpy(d.rec, rec, sizeof(rec));
          dtasks.push_back(d);
        }
      }
      continue;
    }
    {
      Launch L; L.kind = LK_ASSEMBLE_LARGE; L.off = (int64_t)items.size();
      for (int s : large) L.count += asm_items(s, items, 0, nullptr);
      fact_launches.push_back(L);
    }
    int wmax = 0;
    for (int s : large) wmax = std::max(wmax, front_w(s));
    if (lookahead) {   // the first diagonal blocks: nothing to overlap with yet
      Launch D; D.kind = LK_DIAG; D.off = (int64_t)items.size(); D.jb = 0; D.flag = 1;
      for (int s : large) { items.push_back(s); D.count++; }
      fact_launches.push_back(D);
    }
    for (int jb = 0; jb < wmax; jb += NB) {
      Launch T; T.kind = LK_TRSM; T.off = (int64_t)items.size(); T.jb = jb;
      T.flag = lookahead ? 1 : 0;
      T.join_side = (lookahead && jb > 0) ? 1 : 0;
      for (int s : large) {
        int w = front_w(s), m = front_m(s);
        if (w <= jb) continue;
        int nb = std::min(NB, w - jb);
        int nrows = m - (jb + nb);
        // without look-ahead chunk 0 always exists (it owns the diagonal block)
        int nch = (nrows + TRSM_ROWS - 1) / TRSM_ROWS;
        if (!lookahead) nch = std::max(1, nch);
        for (int ch = 0; ch < nch; ch++) { items.push_back(s); items.push_back(ch); T.count++; }
      }
      if (T.count) fact_launches.push_back(T);
      Launch U; U.kind = LK_UPDATE; U.off = (int64_t)items.size(); U.jb = jb; U.mode = 0;
      U.flag = lookahead ? 1 : 0;
      Launch D; D.kind = LK_DIAG; D.jb = jb + NB; D.flag = 0; D.side = 1;
      std::vector<int32_t> dfronts;
      for (int s : large) {
        int w = front_w(s), m = front_m(s);
        if (w <= jb) continue;
        int nb = std::min(NB, w - jb);
        int org = jb + nb;
        if (org >= w) continue;
        dfronts.push_back(s);
        const bool full_next = (w - org) >= NB;   // the next pivot block fills tile (0,0) completely
        int ntj = (w - org + TILE - 1) / TILE, nti = (m - org + TILE - 1) / TILE;
        for (int tj = 0; tj < ntj; tj++)
          for (int ti = tj; ti < nti; ti++) {
            if (lookahead && ti == 0 && tj == 0 && full_next) continue;   // k_diag owns that tile
            items.push_back(s); items.push_back(ti); items.push_back(tj); U.count++;
          }
      }
      if (U.count) fact_launches.push_back(U);
      if (lookahead && !dfronts.empty()) {
        D.off = (int64_t)items.size();
        items.insert(items.end(), dfronts.begin(), dfronts.end());
        D.count = (int)dfronts.size();
        fact_launches.push_back(D);
      }
    }
    {
      Launch U; U.kind = LK_UPDATE; U.off = (int64_t)items.size(); U.mode = 1;
      for (int s : large) {
        int r = front_m(s) - front_w(s);
        int nt = (r + TILE - 1) / TILE;
        for (int tj = 0; tj < nt; tj++)
          for (int ti = tj; ti < nt; ti++) {
            items.push_back(s); items.push_back(ti); items.push_back(tj); U.count++;
          }
      }
      if (U.count) fact_launches.push_back(U);
    }
  }
  // tag every launch with its level and its branch (kernel family): the branches of a level are
  // independent of each other and are captured as parallel branches of the CUDA graph
  fstart.push_back(fact_launches.size()); sstart.push_back(fwd_launches.size()); bstart.push_back(bwd_launches.size());
  auto branch_of = [](const Launch& L) {
    switch (L.kind) {
      case LK_FRONT_TINY: case LK_FWD_TINY: case LK_BWD_TINY: return L.cls;          // 0..3
      case LK_FRONT_SMALL: case LK_FWD: case LK_BWD: return 4 + L.cls;               // 4..7
      case LK_FWD_BIG: case LK_BWD_BIG: return 8;
      default: return 9;                                                              // the tiled path: one ordered chain
    }
  };
  for (int l = 0; l < S.nlevels; l++) {
    for (size_t i = fstart[l]; i < fstart[l + 1]; i++) { fact_launches[i].level = l; fact_launches[i].branch = branch_of(fact_launches[i]); }
    for (size_t i = sstart[l]; i < sstart[l + 1]; i++) { fwd_launches[i].level = l; fwd_launches[i].branch = branch_of(fwd_launches[i]); }
    for (size_t i = bstart[l]; i < bstart[l + 1]; i++) { bwd_launches[i].level = l; bwd_launches[i].branch = branch_of(bwd_launches[i]); }
  }
  // the dataflow launch of the top of the tree: parents' records, absolute counter offsets
  ndag = 0; ndcnt = 0;
  if (!dtasks.empty()) {
    if (dag_sched)
      std::stable_sort(dtasks.begin(), dtasks.end(), [](const DTask& a, const DTask& b) { return a.key < b.key; });
    for (const DTask& d : dtasks) items_dag.insert(items_dag.end(), d.rec, d.rec + 8);
    G.count = (int)dtasks.size();
  }
  if (G.count > 0) {
    const int ndf = (int)dagf.size();
    for (const DagFront& f : dagf) {
      const int p = S.sparent[f.s];
      if (p >= 0) dfr[4 * (size_t)f.df + 2] = dag_of[p];
    }
    plan.dcnt_cb = 0; plan.dcnt_ch = ndf;
    ndcnt = 2 * (int64_t)ndf;
    if ((int64_t)items.size() + (int64_t)items_dag.size() >= (int64_t)INT32_MAX * 4) { snprintf(g_last_error, sizeof(g_last_error), "too many dataflow tasks"); return -1; }
    G.off = (int64_t)items.size();
    items.insert(items.end(), items_dag.begin(), items_dag.end());
    G.level = dag_from_level; G.branch = 9; G.jb = 0; G.mode = 0;
    fact_launches.push_back(G);
    ndag = 1;
  }
  // flat extend-add lists of the fronts that stay in shared memory (k_front_small)
  plan.sf_ptr = nullptr; plan.sf_ent = nullptr;
  if (flat_ok && small_max_m <= 181) {
    std::vector<int64_t> sf_ptr(S.nsuper + 1, 0);
    for (int s = 0; s < S.nsuper; s++) {
      int64_t n = 0;
      const int m = front_m(s);
      if (m <= (int)small_max_m && m > tiny_max_m && !front_dag[s]) {
        n = S.amap_ptr[s + 1] - S.amap_ptr[s];
        for (int q = S.child_ptr[s]; q < S.child_ptr[s + 1]; q++) {
          const int64_t rc = front_m(S.child_idx[q]) - front_w(S.child_idx[q]);
          n += rc * (rc + 1) / 2;
        }
      }
      sf_ptr[s + 1] = sf_ptr[s] + n;
    }
    if (sf_ptr[S.nsuper] > 0 && sf_ptr[S.nsuper] <= flat_small_max) {
      std::vector<int32_t> sf_ent(2 * (size_t)sf_ptr[S.nsuper]);
      struct El { int32_t dest, seq, isA, src; };
      std::vector<El> el;
      for (int s = 0; s < S.nsuper; s++) {
        if (sf_ptr[s + 1] == sf_ptr[s]) continue;
        const int m = front_m(s);
        el.clear();
        for (int64_t q = S.amap_ptr[s]; q < S.amap_ptr[s + 1]; q++) el.push_back({S.amap_pos[q], 0, 1, S.amap_slot[q]});
        for (int q = S.child_ptr[s]; q < S.child_ptr[s + 1]; q++) {
          const int c = S.child_idx[q];
          const int wc = front_w(c), rc = front_m(c) - wc;
          const int32_t* relc = &S.rel[S.rptr[c] + wc];
          for (int j = 0; j < rc; j++)
            for (int i = j; i < rc; i++)
              el.push_back({relc[i] + relc[j] * m, 1 + q - S.child_ptr[s], 0, (int32_t)(S.cbptr[c] + i + (int64_t)j * rc)});
        }
        std::sort(el.begin(), el.end(), [](const El& a, const El& b) { return a.dest != b.dest ? a.dest < b.dest : a.seq < b.seq; });
        int32_t* out = &sf_ent[2 * (size_t)sf_ptr[s]];
        for (size_t e = 0; e < el.size();) {
          size_t r1 = e + 1;
          while (r1 < el.size() && el[r1].dest == el[e].dest && r1 - e < 65535) r1++;
          for (size_t k = e; k < r1; k++) {
            out[2 * k] = el[k].dest | (el[k].isA << 15) | (k == e ? (int32_t)((uint32_t)(r1 - e) << 16) : 0);
            out[2 * k + 1] = el[k].src;
          }
          e = r1;
        }
      }
      if (upload(&d_sf_ptr, sf_ptr, bytes_device)) return -1;
      if (upload(&d_sf_ent, sf_ent, bytes_device)) return -1;
      plan.sf_ptr = d_sf_ptr; plan.sf_ent = d_sf_ent;
    }
  }
  if (upload(&d_dfr, dfr, bytes_device)) return -1;
  {
    // per-tile child lists: counting sort of the (tile, child) pairs by tile (children stay ascending)
    if (tl_tmp.size() >= (size_t)INT32_MAX / 8) { snprintf(g_last_error, sizeof(g_last_error), "too many extend-add pairs"); return -1; }
    std::vector<int32_t> tl_ptr(tl_cnt.size() + 1, 0);
    for (size_t i = 0; i < tl_cnt.size(); i++) tl_ptr[i + 1] = tl_ptr[i] + tl_cnt[i];
    std::vector<int32_t> fill(tl_ptr.begin(), tl_ptr.end() - 1), tl_ent((size_t)DAG_ENT * tl_tmp.size(), 0);
    for (const TlEnt& e : tl_tmp) {
      int32_t* d = &tl_ent[(size_t)DAG_ENT * fill[e.gid]++];
      const int64_t cbo = asm_off[2 * (size_t)e.ce + 1];
      d[0] = asm_rc[e.ce];
      d[1] = (int32_t)(uint32_t)(cbo & 0xffffffffll); d[2] = (int32_t)(cbo >> 32);
      // the child rows [ia, iz) land on rows rel - rlo of the tile, its columns [ja, jz) on columns rel - clo
      uint16_t* rmap = reinterpret_cast<uint16_t*>(d + 8);
      uint16_t* cmap = rmap + 64;
      for (int k = 0; k < 128; k++) rmap[k] = 0xffff;
      const int32_t* relc = &S.rel[asm_off[2 * (size_t)e.ce]];
      for (int i = e.ia; i < e.iz; i++) rmap[relc[i] - e.rlo] = (uint16_t)i;
      for (int j = e.ja; j < e.jz; j++) cmap[relc[j] - e.clo] = (uint16_t)j;
    }
    if (upload(&d_tl_ptr, tl_ptr, bytes_device)) return -1;
    if (upload(&d_tl_ent, tl_ent, bytes_device)) return -1;
    if (fl_ent.size() / 2 >= (size_t)INT32_MAX) { snprintf(g_last_error, sizeof(g_last_error), "too many flat extend-add elements"); return -1; }
    std::vector<int32_t> fl_ptr(fl_cnt.size() + 1, 0);
    for (size_t i = 0; i < fl_cnt.size(); i++) fl_ptr[i + 1] = fl_ptr[i] + fl_cnt[i];
    if (upload(&d_fl_ptr, fl_ptr, bytes_device)) return -1;
    if (upload(&d_fl_ent, fl_ent, bytes_device)) return -1;
    plan.dfr = d_dfr; plan.tl_ptr = d_tl_ptr; plan.tl_ent = d_tl_ent; plan.fl_ptr = d_fl_ptr; plan.fl_ent = d_fl_ent;
  }
  // staging area of the factored diagonal blocks + the write-back launch that ends a factorization
  {
    std::vector<int64_t> dsptr(S.nsuper + 1, 0);
    Launch W; W.kind = LK_DIAG_WRITEBACK; W.off = (int64_t)items.size();
    int64_t off = 0;
    for (int s = 0; s < S.nsuper; s++) {
      dsptr[s] = off;
      if (front_m(s) <= (int)small_max_m && !front_dag[s]) continue;
      int nblk = (front_w(s) + NB - 1) / NB;
      if (!front_dag[s])   // k_front_dag writes the diagonal blocks in place
        for (int bi = 0; bi < nblk; bi++) { items.push_back(s); items.push_back(bi); W.count++; }
      off += (int64_t)nblk * NB * NB;
    }
    dsptr[S.nsuper] = off;
    W.level = S.nlevels; W.branch = 9;
    if (W.count) fact_launches.push_back(W);
    // explicit inverses of the diagonal blocks of the chained big-front solves (after the write-back:
    // the launch-chain fronts get their factored diagonal blocks from the staging area)
    {
      std::vector<int32_t> linv_idx(S.nsuper, -1);
      Launch V; V.kind = LK_LINV; V.off = (int64_t)items.size();
      n_linv = 0;
      for (int s = 0; s < S.nsuper; s++) {
        const int nblk = (front_w(s) + SB - 1) / SB;
        if (front_m(s) <= (int)solve_big_m || nblk < inv_min_blk || inv_min_blk <= 0) continue;
        linv_idx[s] = (int32_t)n_linv;
        for (int c = 0; c < nblk; c++) { items.push_back(s); items.push_back(c); V.count++; }
        n_linv += nblk;
      }
      V.level = S.nlevels + 1; V.branch = 9;
      if (V.count) fact_launches.push_back(V);
      if (upload(&d_linv_idx, linv_idx, bytes_device)) return -1;
      if (dalloc(&d_linv, (size_t)std::max<int64_t>(n_linv, 1) * SB * SB, bytes_device)) return -1;
      plan.linv_idx = d_linv_idx; plan.Linv = d_linv;
    }
    if (upload(&d_dsptr, dsptr, bytes_device)) return -1;
    if (dalloc(&d_dstage, (size_t)off, bytes_device)) return -1;
    plan.dstage = d_dstage; plan.dsptr = d_dsptr;
  }
  if (upload(&d_asm_cptr, asm_cptr, bytes_device)) return -1;
  if (upload(&d_asm_ent, asm_ent, bytes_device)) return -1;
  if (upload(&d_asm_rc, asm_rc, bytes_device)) return -1;
  if ((int64_t)asm_cptr.size() >= (int64_t)INT32_MAX) { snprintf(g_last_error, sizeof(g_last_error), "too many tiled-front columns"); return -1; }
  if (upload(&d_asm_off, asm_off, bytes_device)) return -1;
  plan.asm_cptr = d_asm_cptr; plan.asm_ent = d_asm_ent; plan.asm_rc = d_asm_rc; plan.asm_off = d_asm_off;
  {
    // forward solve of the fronts that are not "big": per front row, the child update entries that
    // land on it, children ascending = the order the per-child loop added them in (bit-identical
    // sums).  One pass replaces, per child, a chain of dependent loads (child -> scol / rptr / uptr ->
    // rel -> upd) and a barrier: the gather of a front is now two load latencies whatever its
    // number of children.
    const int64_t nrows = S.rptr[S.nsuper];
    if (nrows >= (int64_t)INT32_MAX) { snprintf(g_last_error, sizeof(g_last_error), "front row storage exceeds int32 offsets"); return -1; }
    std::vector<int32_t> ug_ptr((size_t)nrows + 1, 0), ug_src;
    std::vector<int32_t> cnt;
    for (int s = 0; s < S.nsuper; s++) {
      const int m = front_m(s);
      const int64_t r0 = S.rptr[s];
      if (m > (int)solve_big_m) continue;      // k_fwd_big has its own gather (sb_ptr / sb_src)
      for (int q = S.child_ptr[s]; q < S.child_ptr[s + 1]; q++) {
        const int c = S.child_idx[q];
        const int wc = front_w(c), rc = front_m(c) - wc;
        const int32_t* relc = &S.rel[S.rptr[c] + wc];
        for (int k = 0; k < rc; k++) ug_ptr[r0 + relc[k] + 1]++;
      }
    }
    for (int64_t i = 0; i < nrows; i++) ug_ptr[i + 1] += ug_ptr[i];
    ug_src.resize((size_t)ug_ptr[nrows]);
    std::vector<uint8_t> ug_row(ug_src.size() + 1, 0);
    for (int s = 0; s < S.nsuper; s++) {
      const int m = front_m(s);
      const int64_t r0 = S.rptr[s];
      if (m > (int)solve_big_m) continue;
      cnt.assign(m, 0);
      for (int q = S.child_ptr[s]; q < S.child_ptr[s + 1]; q++) {
        const int c = S.child_idx[q];
        const int wc = front_w(c), rc = front_m(c) - wc;
        const int32_t* relc = &S.rel[S.rptr[c] + wc];
        for (int k = 0; k < rc; k++) {
          const size_t e = (size_t)ug_ptr[r0 + relc[k]] + cnt[relc[k]]++;
          ug_src[e] = (int32_t)(S.uptr[c] + k);
          ug_row[e] = (uint8_t)std::min(relc[k], 255);
        }
      }
    }
    if (upload(&d_ug_ptr, ug_ptr, bytes_device)) return -1;
    if (upload(&d_ug_src, ug_src, bytes_device)) return -1;
    if (upload(&d_ug_row, ug_row, bytes_device)) return -1;
    plan.ug_ptr = d_ug_ptr; plan.ug_src = d_ug_src; plan.ug_row = d_ug_row;
  }
  if (upload(&d_sb_ptr, sb_ptr, bytes_device)) return -1;
  if (upload(&d_sb_src, sb_src, bytes_device)) return -1;
  if (upload(&d_sb_flag, sb_flag, bytes_device)) return -1;
  if (dalloc(&d_ypub, (size_t)(2 * S.N), bytes_device)) return -1;   // forward | backward publication slots
  if (dalloc(&d_stk, (size_t)std::max(n_stk, 1), bytes_device)) return -1;
  plan.sb_ptr = d_sb_ptr; plan.sb_src = d_sb_src; plan.sb_flag = d_sb_flag;
  if (dalloc(&d_tflag, (size_t)(ntflag + 1 + ndcnt), bytes_device)) return -1;   // tile flags | ticket counter | counters
  plan.dcnt = d_tflag + ntflag + 1;
  std::reverse(bwd_launches.begin(), bwd_launches.end());
  if (upload(&d_items, items, bytes_device)) return -1;
  return 0;
}

int Engine::init(int dev) {
  const double t0 = wall();
  device = dev;
  B2_CUDA_OK(cudaSetDevice(dev));
  B2_CUDA_OK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  for (auto& e : ev) B2_CUDA_OK(cudaEventCreate(&e));
  for (auto& e : tev) B2_CUDA_OK(cudaEventCreate(&e));
  for (auto& s : bstream) B2_CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  B2_CUDA_OK(cudaStreamCreateWithFlags(&sstream, cudaStreamNonBlocking));
  B2_CUDA_OK(cudaEventCreateWithFlags(&ev_sfork, cudaEventDisableTiming));
  B2_CUDA_OK(cudaEventCreateWithFlags(&ev_sjoin, cudaEventDisableTiming));
  B2_CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  for (auto& e : ev_join) B2_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  const Symbolic& S = sym;
  if ((size_t)(S.max_front + SNB) * sizeof(double) > 200 * 1024) {
    snprintf(g_last_error, sizeof(g_last_error), "front of order %d exceeds the solve kernels' shared memory", S.max_front);
    return -1;
  }
  std::vector<int32_t> slot_ptr32(S.slot_ptr.begin(), S.slot_ptr.end());
  if (upload(&d_slot_ptr, slot_ptr32, bytes_device)) return -1;
  if (upload(&d_coo_sorted, S.coo_sorted, bytes_device)) return -1;
  if (dalloc(&d_vals, (size_t)S.nnz, bytes_device)) return -1;
  if (dalloc(&d_nzval, (size_t)S.nnzA, bytes_device)) return -1;
  if (upload(&d_rho_slot, S.rho_slot, bytes_device)) return -1;
  if (upload(&d_delta_slot, S.delta_slot, bytes_device)) return -1;
  if (dalloc(&d_rho_base, (size_t)S.nvar, bytes_device)) return -1;
  if (dalloc(&d_delta_base, (size_t)S.ncon, bytes_device)) return -1;
  if (upload(&d_scol, S.scol, bytes_device)) return -1;
  if (upload(&d_rptr, S.rptr, bytes_device)) return -1;
  if (upload(&d_lptr, S.lptr, bytes_device)) return -1;
  if (upload(&d_cbptr, S.cbptr, bytes_device)) return -1;
  if (upload(&d_uptr, S.uptr, bytes_device)) return -1;
  if (upload(&d_rowidx, S.rowidx, bytes_device)) return -1;
  if (upload(&d_rel, S.rel, bytes_device)) return -1;
  if (upload(&d_child_ptr, S.child_ptr, bytes_device)) return -1;
  if (upload(&d_child_idx, S.child_idx, bytes_device)) return -1;
  if (upload(&d_amap_ptr, S.amap_ptr, bytes_device)) return -1;
  if (upload(&d_amap_slot, S.amap_slot, bytes_device)) return -1;
  if (upload(&d_amap_pos, S.amap_pos, bytes_device)) return -1;
  if (upload(&d_perm, S.perm, bytes_device)) return -1;
  if (upload(&d_Sp, S.Sp, bytes_device)) return -1;
  if (upload(&d_Sj, S.Sj, bytes_device)) return -1;
  if (upload(&d_Sslot, S.Sslot, bytes_device)) return -1;
  if (dalloc(&d_Lx, (size_t)S.nnzL_store + 8, bytes_device)) return -1;   // + 8: the bulk copies of k_update read up to two doubles past a tile
  if (dalloc(&d_CB, (size_t)S.cb_store, bytes_device)) return -1;
  if (dalloc(&d_dvec, (size_t)S.N, bytes_device)) return -1;
  if (dalloc(&d_counts, 8, bytes_device)) return -1;
  d_flags = reinterpret_cast<int*>(d_counts + 4);
  if (dalloc(&d_x, (size_t)S.N, bytes_device)) return -1;
  if (dalloc(&d_upd, (size_t)S.uptr[S.nsuper], bytes_device)) return -1;
  if (dalloc(&d_rhs, (size_t)S.N, bytes_device)) return -1;
  if (dalloc(&d_sol, (size_t)S.N, bytes_device)) return -1;
  if (dalloc(&d_res, (size_t)S.N, bytes_device)) return -1;
  if (dalloc(&d_out, (size_t)S.N, bytes_device)) return -1;
  if (dalloc(&d_part, 1024 + 8, bytes_device)) return -1;
  B2_CUDA_OK(cudaMallocHost((void**)&h_counts, 8 * sizeof(unsigned long long)));
  B2_CUDA_OK(cudaMallocHost((void**)&h_scalars, 8 * sizeof(double)));
  plan.scol = d_scol; plan.rptr = d_rptr; plan.lptr = d_lptr; plan.cbptr = d_cbptr; plan.uptr = d_uptr;
  plan.rowidx = d_rowidx; plan.rel = d_rel; plan.child_ptr = d_child_ptr; plan.child_idx = d_child_idx;
  plan.amap_ptr = d_amap_ptr; plan.amap_slot = d_amap_slot; plan.amap_pos = d_amap_pos;
  plan.nzval = d_nzval; plan.Lx = d_Lx; plan.CB = d_CB; plan.dvec = d_dvec; plan.flags = d_flags;
  {
    cudaDeviceProp prop;
    B2_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    dag_ctas = 2 * prop.multiProcessorCount;   // k_front_dag: __launch_bounds__(256, 2), DAG_SMEM fits twice
  }
  if (build_plan()) return -1;
  const int big = 200 * 1024;
  B2_CUDA_OK(cudaFuncSetAttribute(k_front_small<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  B2_CUDA_OK(cudaFuncSetAttribute(k_front_small<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  B2_CUDA_OK(cudaFuncSetAttribute(k_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM));
  B2_CUDA_OK(cudaFuncSetAttribute(k_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
  B2_CUDA_OK(cudaFuncSetAttribute(k_front_dag, cudaFuncAttributeMaxDynamicSharedMemorySize, DAG_SMEM_EXCL));
  B2_CUDA_OK(cudaFuncSetAttribute(k_bwd_big, cudaFuncAttributeMaxDynamicSharedMemorySize, big - 48 * 1024));
  B2_CUDA_OK(cudaFuncSetAttribute(k_fwd<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  B2_CUDA_OK(cudaFuncSetAttribute(k_bwd<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  B2_CUDA_OK(cudaDeviceSynchronize());
  t_plan = wall() - t0;
  return 0;
}

void Engine::destroy() {
#ifndef B2_EMULATE
  if (g_fact) cudaGraphExecDestroy(g_fact);
  if (g_fwdbwd) cudaGraphExecDestroy(g_fwdbwd);
#endif
  for (void* p : registered) cudaHostUnregister(p);
  void* ptrs[] = {d_slot_ptr, d_coo_sorted, d_vals, d_nzval, d_rho_slot, d_delta_slot, d_rho_base,
                  d_delta_base, d_scol, d_rowidx, d_rel, d_child_ptr, d_child_idx, d_amap_slot,
                  d_amap_pos, d_perm, d_rptr, d_lptr, d_cbptr, d_uptr, d_amap_ptr, d_Lx, d_CB, d_dvec,
                  d_counts, d_items, d_dstage, d_dsptr, d_asm_cptr, d_asm_ent, d_asm_rc, d_asm_off, d_sb_ptr, d_sb_src, d_sb_flag, d_linv_idx, d_linv, d_ug_ptr, d_ug_src, d_ug_row, d_ypub, d_stk, d_tflag, d_dfr, d_tl_ptr, d_tl_ent, d_fl_ptr, d_fl_ent, d_sf_ptr, d_sf_ent, d_vals2, d_mismatch, d_x, d_upd, d_rhs, d_sol, d_res, d_out, d_part, d_Sp, d_Sj, d_Sslot};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (h_mismatch) cudaFreeHost(h_mismatch);
  if (cstream) cudaStreamDestroy(cstream);
  if (ev_cmp) cudaEventDestroy(ev_cmp);
  if (ev_cfork) cudaEventDestroy(ev_cfork);
  if (h_counts) cudaFreeHost(h_counts);
  if (h_scalars) cudaFreeHost(h_scalars);
  for (auto& e : ev) if (e) cudaEventDestroy(e);
  for (auto& e : tev) if (e) cudaEventDestroy(e);
  if (ev_sfork) cudaEventDestroy(ev_sfork);
  if (ev_sjoin) cudaEventDestroy(ev_sjoin);
  if (sstream) cudaStreamDestroy(sstream);
  if (ev_fork) cudaEventDestroy(ev_fork);
  for (auto& e : ev_join) if (e) cudaEventDestroy(e);
  for (auto& s : bstream) if (s) cudaStreamDestroy(s);
  if (stream) cudaStreamDestroy(stream);
}

int Engine::launch_one(const Launch& L, cudaStream_t st) {
  const int32_t* it = d_items + L.off;
  switch (L.kind) {
    case LK_FRONT_SMALL:
      if (L.cls == 0) { auto kfn = k_front_small<32, FPB32>; B2_LAUNCH(kfn, (L.count + FPB32 - 1) / FPB32, 32 * FPB32, L.smem * FPB32, st, plan, it, L.count, L.smem); }
      else if (L.cls == 1) { auto kfn = k_front_small<64, 1>; B2_LAUNCH(kfn, L.count, 64, L.smem, st, plan, it, L.count, 0); }
      else if (L.cls == 2) { auto kfn = k_front_small<128, 1>; B2_LAUNCH(kfn, L.count, 128, L.smem, st, plan, it, L.count, 0); }
      else { auto kfn = k_front_small<256, 1>; B2_LAUNCH(kfn, L.count, 256, L.smem, st, plan, it, L.count, 0); }
      break;
    case LK_ASSEMBLE_LARGE:
      B2_LAUNCH(k_assemble_large, L.count, 256, 0, st, plan, it, L.count);
      break;
    case LK_TRSM:
      B2_LAUNCH(k_trsm, L.count, TRSM_THREADS, TRSM_SMEM, st, plan, it, L.count, L.jb, L.flag);
      break;
    case LK_DIAG:
      B2_LAUNCH(k_diag, L.count, 256, DIAG_SMEM, st, plan, it, L.count, L.jb, L.flag);
      break;
    case LK_LINV:
      B2_LAUNCH(k_linv, L.count, 256, 0, st, plan, it, L.count);
      break;
    case LK_DIAG_WRITEBACK:
      B2_LAUNCH(k_diag_writeback, L.count, 256, 0, st, plan, it, L.count);
      break;
    case LK_UPDATE:
      if (update_tma) B2_LAUNCH(k_update<true>, L.count, 256, 0, st, plan, it, L.count, L.jb, NB, L.mode, L.flag);
      else B2_LAUNCH(k_update<false>, L.count, 256, 0, st, plan, it, L.count, L.jb, NB, L.mode, L.flag);
      break;
    case LK_DAG:
      // few tasks (the top of the tree): one CTA per SM -- asking for more than half of the shared
      // memory keeps a second CTA off the SM, so the pivot chain does not share its issue slots and
      // FP64 pipe with a neighbour's update (measured: LDL^T of a 64 x 64 tile 16 -> 11.5 us)
      if (L.count <= dag_excl_max)
        B2_LAUNCH(k_front_dag, std::min(L.count, dag_ctas / 2), 256, DAG_SMEM_EXCL, st, plan, it, L.count, d_tflag,
                  d_tflag + ntflag, L.mode);
      else
        B2_LAUNCH(k_front_dag, std::min(L.count, dag_ctas), 256, DAG_SMEM, st, plan, it, L.count, d_tflag,
                  d_tflag + ntflag, L.mode);
      break;
    case LK_FWD:
      if (L.cls == 0) { auto kfn = k_fwd<32, FPB32>; B2_LAUNCH(kfn, (L.count + FPB32 - 1) / FPB32, 32 * FPB32, L.smem * FPB32, st, plan, it, L.count, d_x, d_upd, L.smem); }
      else if (L.cls == 1) { auto kfn = k_fwd<64, 1>; B2_LAUNCH(kfn, L.count, 64, L.smem, st, plan, it, L.count, d_x, d_upd, 0); }
      else if (L.cls == 2) { auto kfn = k_fwd<128, 1>; B2_LAUNCH(kfn, L.count, 128, L.smem, st, plan, it, L.count, d_x, d_upd, 0); }
      else { auto kfn = k_fwd<256, 1>; B2_LAUNCH(kfn, L.count, 256, L.smem, st, plan, it, L.count, d_x, d_upd, 0); }
      break;
    case LK_BWD:
      if (L.cls == 0) { auto kfn = k_bwd<32, FPB32>; B2_LAUNCH(kfn, (L.count + FPB32 - 1) / FPB32, 32 * FPB32, L.smem * FPB32, st, plan, it, L.count, d_x, L.smem); }
      else if (L.cls == 1) { auto kfn = k_bwd<64, 1>; B2_LAUNCH(kfn, L.count, 64, L.smem, st, plan, it, L.count, d_x, 0); }
      else if (L.cls == 2) { auto kfn = k_bwd<128, 1>; B2_LAUNCH(kfn, L.count, 128, L.smem, st, plan, it, L.count, d_x, 0); }
      else { auto kfn = k_bwd<256, 1>; B2_LAUNCH(kfn, L.count, 256, L.smem, st, plan, it, L.count, d_x, 0); }
      break;
    case LK_FRONT_TINY:
      if (L.cls == 0) B2_LAUNCH(k_front_tiny<4>, (L.count + tiny_nt(4) - 1) / tiny_nt(4), tiny_nt(4), 0, st, plan, it, L.count);
      else B2_LAUNCH(k_front_tiny<8>, (L.count + tiny_nt(8) - 1) / tiny_nt(8), tiny_nt(8), 0, st, plan, it, L.count);
      break;
    case LK_FWD_TINY:
      if (L.cls == 0) B2_LAUNCH(k_fwd_tiny<4>, (L.count + tiny_nt(4) - 1) / tiny_nt(4), tiny_nt(4), 0, st, plan, it, L.count, d_x, d_upd);
      else if (L.cls == 1) B2_LAUNCH(k_fwd_tiny<8>, (L.count + tiny_nt(8) - 1) / tiny_nt(8), tiny_nt(8), 0, st, plan, it, L.count, d_x, d_upd);
      else if (L.cls == 2) B2_LAUNCH(k_fwd_tiny<16>, (L.count + tiny_nt(16) - 1) / tiny_nt(16), tiny_nt(16), 0, st, plan, it, L.count, d_x, d_upd);
      else B2_LAUNCH(k_fwd_tiny<32>, (L.count + tiny_nt(32) - 1) / tiny_nt(32), tiny_nt(32), 0, st, plan, it, L.count, d_x, d_upd);
      break;
    case LK_BWD_TINY:
      if (L.cls == 0) B2_LAUNCH(k_bwd_tiny<4>, (L.count + tiny_nt(4) - 1) / tiny_nt(4), tiny_nt(4), 0, st, plan, it, L.count, d_x);
      else if (L.cls == 1) B2_LAUNCH(k_bwd_tiny<8>, (L.count + tiny_nt(8) - 1) / tiny_nt(8), tiny_nt(8), 0, st, plan, it, L.count, d_x);
      else if (L.cls == 2) B2_LAUNCH(k_bwd_tiny<16>, (L.count + tiny_nt(16) - 1) / tiny_nt(16), tiny_nt(16), 0, st, plan, it, L.count, d_x);
      else B2_LAUNCH(k_bwd_tiny<32>, (L.count + tiny_nt(32) - 1) / tiny_nt(32), tiny_nt(32), 0, st, plan, it, L.count, d_x);
      break;
    case LK_FWD_BIG:
      B2_LAUNCH(k_fwd_big, L.count, 256, 0, st, plan, it, L.count, d_x, d_upd, d_ypub, d_stk + L.jb);
      break;
    case LK_BWD_BIG:
      B2_LAUNCH(k_bwd_big, L.count, 256, L.smem, st, plan, it, L.count, d_x, d_ypub + sym.N, d_stk + L.jb);
      break;
    default: break;
  }
  return 0;
}

// Launch a list level by level.  Inside a level the launches of different branches (kernel
// families) are independent: they go to side streams forked from / joined back into the main
// stream, so that under graph capture they become parallel branches of the CUDA graph.
int Engine::run_list(const std::vector<Launch>& LL, bool allow_fork) {
  size_t i = 0;
  while (i < LL.size()) {
    size_t j = i;
    unsigned mask = 0;
    while (j < LL.size() && LL[j].level == LL[i].level) { mask |= 1u << LL[j].branch; j++; }
    // A level that contains a flag-chained multi-CTA solve (branch 8) is launched in order on the
    // main stream: its CTAs form a serial chain, co-resident kernels slow exactly the critical
    // path and the fork/join edges add latency to it (measured: +6..10 % on the C4 solve).
    const unsigned excl = mask & (1u << 8);
    mask &= ~(1u << 8);
    const bool fork = use_branches && allow_fork && !excl && (mask & (mask - 1)) != 0;
    // a launch of the tiled chain may go to the side stream (look-ahead diagonal factorization) and
    // a later one waits for it
    bool side_pending = false;
    auto chain_launch = [&](const Launch& L, cudaStream_t cs) -> int {
      if (L.join_side && side_pending) {
        B2_CUDA_OK(cudaStreamWaitEvent(cs, ev_sjoin, 0));
        side_pending = false;
      }
      if (L.side && use_branches) {
        B2_CUDA_OK(cudaEventRecord(ev_sfork, cs));
        B2_CUDA_OK(cudaStreamWaitEvent(sstream, ev_sfork, 0));
        launch_one(L, sstream);
        B2_CUDA_OK(cudaEventRecord(ev_sjoin, sstream));
        side_pending = true;
      } else {
        launch_one(L, cs);
      }
      return 0;
    };
    if (!fork) {
      for (size_t q = i; q < j; q++)
        if (LL[q].branch != 8 && chain_launch(LL[q], stream)) return -1;
      if (side_pending) { B2_CUDA_OK(cudaStreamWaitEvent(stream, ev_sjoin, 0)); side_pending = false; }
    } else {
      B2_CUDA_OK(cudaEventRecord(ev_fork, stream));
      for (int b = 0; b < NBRANCH; b++)
        if (mask & (1u << b)) B2_CUDA_OK(cudaStreamWaitEvent(bstream[b], ev_fork, 0));
      for (size_t q = i; q < j; q++)
        if (LL[q].branch != 8 && chain_launch(LL[q], bstream[LL[q].branch])) return -1;
      if (side_pending) { B2_CUDA_OK(cudaStreamWaitEvent(bstream[9], ev_sjoin, 0)); side_pending = false; }
      for (int b = 0; b < NBRANCH; b++)
        if (mask & (1u << b)) {
          B2_CUDA_OK(cudaEventRecord(ev_join[b], bstream[b]));
          B2_CUDA_OK(cudaStreamWaitEvent(stream, ev_join[b], 0));
        }
    }
    if (excl)
      for (size_t q = i; q < j; q++)
        if (LL[q].branch == 8) launch_one(LL[q], stream);
    i = j;
  }
  B2_CUDA_OK(cudaGetLastError());
  return 0;
}

int Engine::run_factor_launches() {
  // tile flags and ticket counters of the dataflow launches
  if (ndag > 0) B2_CUDA_OK(cudaMemsetAsync(d_tflag, 0, (size_t)(ntflag + 1 + ndcnt) * sizeof(int), stream));
  return run_list(fact_launches, true);
}

int Engine::run_solve_launches() {
  // publication slots of the multi-CTA solves: all-ones = "not yet published" (poll_value)
  if (nsflag > 0) B2_CUDA_OK(cudaMemsetAsync(d_ypub, 0xFF, (size_t)(2 * sym.N) * sizeof(double), stream));
  if (n_stk > 0) B2_CUDA_OK(cudaMemsetAsync(d_stk, 0, (size_t)n_stk * sizeof(int), stream));
  // measured: on systems with big fronts (C4) forking the solve levels costs more than it gains
  // (3.14 -> 3.49 ms), on systems made of small fronts only (C2) it gains 25 %
  const bool fork = solve_fork < 0 ? nsflag == 0 : solve_fork != 0;
  if (run_list(fwd_launches, fork)) return -1;
  return run_list(bwd_launches, fork);
}

#ifdef B2_TIMING
extern "C" int b2_debug_clocks(long long* out64) {
  return cudaMemcpyFromSymbol(out64, b2_dbg, 64 * sizeof(long long)) == cudaSuccess ? 0 : -1;
}
extern "C" int b2_debug_dag_trace(long long* out, long long nlongs) {
  return cudaMemcpyFromSymbol(out, b2_dag_trace, (size_t)nlongs * sizeof(long long)) == cudaSuccess ? 0 : -1;
}
#endif

// Developer aid: replay the factorization (which = 0) or one forward+backward sweep (which = 1)
// launch by launch, outside the CUDA graph, with an event after every launch: warm-cache
// per-launch device times (ncu's are cold-cache).  Needs a previous factorize / solve.
int Engine::profile(int which, int max, int* kinds, int* cls, int* counts, double* ms, int* n) {
  if (!have_vals) { snprintf(g_last_error, sizeof(g_last_error), "b2_profile before any factorization"); return -2; }
  B2_CUDA_OK(cudaSetDevice(device));
  std::vector<const Launch*> LL;
  if (which == 0) for (const Launch& L : fact_launches) LL.push_back(&L);
  else { for (const Launch& L : fwd_launches) LL.push_back(&L); for (const Launch& L : bwd_launches) LL.push_back(&L); }
  std::vector<cudaEvent_t> evs(LL.size() + 1);
  for (auto& e : evs) B2_CUDA_OK(cudaEventCreate(&e));
  for (int rep = 0; rep < 2; rep++) {   // second pass is the warm one
    if (which == 0) {
      B2_CUDA_OK(cudaMemsetAsync(d_counts, 0, 8 * sizeof(unsigned long long), stream));
      if (ndag > 0) B2_CUDA_OK(cudaMemsetAsync(d_tflag, 0, (size_t)(ntflag + 1 + ndcnt) * sizeof(int), stream));
    } else if (nsflag > 0) {
      B2_CUDA_OK(cudaMemsetAsync(d_ypub, 0xFF, (size_t)(2 * sym.N) * sizeof(double), stream));
      B2_CUDA_OK(cudaMemsetAsync(d_stk, 0, (size_t)std::max(n_stk, 1) * sizeof(int), stream));
    }
    B2_CUDA_OK(cudaEventRecord(evs[0], stream));
    for (size_t i = 0; i < LL.size(); i++) {
      launch_one(*LL[i], stream);
      B2_CUDA_OK(cudaEventRecord(evs[i + 1], stream));
    }
    B2_CUDA_OK(cudaStreamSynchronize(stream));
  }
  B2_CUDA_OK(cudaGetLastError());
  int cnt = 0;
  for (size_t i = 0; i < LL.size() && cnt < max; i++, cnt++) {
    float f = 0;
    cudaEventElapsedTime(&f, evs[i], evs[i + 1]);
    kinds[cnt] = LL[i]->kind; cls[cnt] = LL[i]->kind == LK_UPDATE ? LL[i]->mode : LL[i]->cls;
    counts[cnt] = LL[i]->count; ms[cnt] = f;
  }
  *n = cnt;
  for (auto& e : evs) cudaEventDestroy(e);
  return 0;
}

int Engine::assemble_and_factor(double eig_tol, int64_t* npos, int64_t* nzero, int64_t* nneg,
                                int* breakdown, bool do_assemble) {
  const Symbolic& S = sym;
  B2_CUDA_OK(cudaEventRecord(ev[1], stream));
  if (do_assemble) {
    const int nb = (int)((S.nnzA + 255) / 256);
    B2_LAUNCH(k_assemble_csc, nb, 256, 0, stream, S.nnzA, d_slot_ptr, d_coo_sorted, d_vals, d_nzval);
    if (S.shift_ok) {
      if (S.nvar > 0)
        B2_LAUNCH(k_diag_base, (int)((S.nvar + 255) / 256), 256, 0, stream, (int)S.nvar, d_rho_slot,
                  d_slot_ptr, d_coo_sorted, d_vals, d_rho_base);
      if (S.ncon > 0)
        B2_LAUNCH(k_diag_base, (int)((S.ncon + 255) / 256), 256, 0, stream, (int)S.ncon, d_delta_slot,
                  d_slot_ptr, d_coo_sorted, d_vals, d_delta_base);
    }
  }
  B2_CUDA_OK(cudaEventRecord(ev[2], stream));
  B2_CUDA_OK(cudaMemsetAsync(d_counts, 0, 8 * sizeof(unsigned long long), stream));
#ifndef B2_EMULATE
  if (use_graph) {
    if (!g_fact) {
      cudaGraph_t g = nullptr;
      B2_CUDA_OK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
      int rc = run_factor_launches();
      cudaError_t ce = cudaStreamEndCapture(stream, &g);
      if (rc || ce != cudaSuccess) {   // (the capture has been ended either way: the stream stays usable)
        if (!rc)    // a failed launch inside the capture has already left its own message
          snprintf(g_last_error, sizeof(g_last_error), "graph capture of the factorization failed: %s",
                   cudaGetErrorString(ce));
        if (ce == cudaSuccess && g) cudaGraphDestroy(g);
        return -1;
      }
      B2_CUDA_OK(cudaGraphInstantiate(&g_fact, g, 0));
      cudaGraphDestroy(g);
    }
    B2_CUDA_OK(cudaGraphLaunch(g_fact, stream));
  } else
#endif
  {
    if (run_factor_launches()) return -1;
  }
  {
    int nb = (int)std::min<int64_t>((S.N + 255) / 256, 1184);
    B2_LAUNCH(k_inertia, nb, 256, 0, stream, d_dvec, S.N, eig_tol, d_counts);
  }
  B2_CUDA_OK(cudaEventRecord(ev[3], stream));
  B2_CUDA_OK(cudaMemcpyAsync(h_counts, d_counts, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream));
  B2_CUDA_OK(cudaStreamSynchronize(stream));
  B2_CUDA_OK(cudaGetLastError());
  if (npos) *npos = (int64_t)h_counts[0];
  if (nzero) *nzero = (int64_t)h_counts[1];
  if (nneg) *nneg = (int64_t)h_counts[2];
  if (breakdown) *breakdown = (int)(h_counts[4] & 0xffffffffull) != 0;
  factored = true;
  float ms = 0;
  cudaEventElapsedTime(&ms, ev[0], ev[1]); last_ms[0] = ms;
  cudaEventElapsedTime(&ms, ev[1], ev[2]); last_ms[1] = ms;
  cudaEventElapsedTime(&ms, ev[2], ev[3]); last_ms[2] = ms;
  return 0;
}

int Engine::factorize_host(const double* vals, double eig_tol, int64_t* npos, int64_t* nzero,
                           int64_t* nneg, int* breakdown) {
  B2_CUDA_OK(cudaSetDevice(device));
  B2_CUDA_OK(cudaEventRecord(ev[0], stream));
  B2_CUDA_OK(cudaMemcpyAsync(d_vals, vals, (size_t)sym.nnz * sizeof(double), cudaMemcpyHostToDevice, stream));
  have_vals = true;
  return assemble_and_factor(eig_tol, npos, nzero, nneg, breakdown, true);
}

int Engine::factorize_dev(const double* d_vals_in, double eig_tol, int64_t* npos, int64_t* nzero,
                          int64_t* nneg, int* breakdown) {
  B2_CUDA_OK(cudaSetDevice(device));
  B2_CUDA_OK(cudaEventRecord(ev[0], stream));
  if (d_vals_in != d_vals)
    B2_CUDA_OK(cudaMemcpyAsync(d_vals, d_vals_in, (size_t)sym.nnz * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  have_vals = true;
  return assemble_and_factor(eig_tol, npos, nzero, nneg, breakdown, true);
}

int Engine::refactorize_shift(double rho, double delta, double eig_tol, int64_t* npos, int64_t* nzero,
                              int64_t* nneg, int* breakdown) {
  const Symbolic& S = sym;
  if (!S.shift_ok) {
    snprintf(g_last_error, sizeof(g_last_error), "COO layout has no canonical rho/delta segments; re-upload with b2_factorize");
    return -2;
  }
  if (!have_vals) {
    snprintf(g_last_error, sizeof(g_last_error), "b2_refactorize_shift before any b2_factorize");
    return -2;
  }
  B2_CUDA_OK(cudaSetDevice(device));
  B2_CUDA_OK(cudaEventRecord(ev[0], stream));
  if (S.nvar > 0)
    B2_LAUNCH(k_diag_shift, (int)((S.nvar + 255) / 256), 256, 0, stream, (int)S.nvar, d_rho_slot, d_rho_base, rho, d_nzval);
  if (S.ncon > 0 && delta == delta)
    B2_LAUNCH(k_diag_shift, (int)((S.ncon + 255) / 256), 256, 0, stream, (int)S.ncon, d_delta_slot, d_delta_base, -delta, d_nzval);
  return assemble_and_factor(eig_tol, npos, nzero, nneg, breakdown, false);
}

// A rho retry that cannot go wrong: same result as factorize_host(vals), bit for bit, whatever the
// caller did to vals.  The caller believes that only the trailing rho segment changed since the last
// upload and now holds the constant `rho` (the protocol of newton_system!, src/CaNNOLeS.jl:1029-1043):
// the shifted matrix is factorized right away while, on a copy stream, vals is uploaded into the
// second buffer and compared on the device with the previous upload.  If the belief was wrong the
// factorization is redone from the fresh upload (assembly included).
int Engine::factorize_retry(const double* vals, double rho, double eig_tol, int64_t* npos, int64_t* nzero,
                            int64_t* nneg, int* breakdown, int* speculation_held) {
  const Symbolic& S = sym;
  if (speculation_held) *speculation_held = 0;
  if (!S.shift_ok || !have_vals || S.nvar == 0) return factorize_host(vals, eig_tol, npos, nzero, nneg, breakdown);
  B2_CUDA_OK(cudaSetDevice(device));
  if (!d_vals2) {
    B2_CUDA_OK(cudaMalloc((void**)&d_vals2, (size_t)S.nnz * sizeof(double)));
    B2_CUDA_OK(cudaMalloc((void**)&d_mismatch, sizeof(int)));
    B2_CUDA_OK(cudaMallocHost((void**)&h_mismatch, sizeof(int)));
    B2_CUDA_OK(cudaStreamCreateWithFlags(&cstream, cudaStreamNonBlocking));
    B2_CUDA_OK(cudaEventCreateWithFlags(&ev_cmp, cudaEventDisableTiming));
    B2_CUDA_OK(cudaEventCreateWithFlags(&ev_cfork, cudaEventDisableTiming));
    bytes_device += (double)S.nnz * sizeof(double);
  }
  // the copy stream starts after whatever the main stream still does with d_vals (nothing reads d_vals2)
  B2_CUDA_OK(cudaEventRecord(ev_cfork, stream));
  B2_CUDA_OK(cudaStreamWaitEvent(cstream, ev_cfork, 0));
  B2_CUDA_OK(cudaMemsetAsync(d_mismatch, 0, sizeof(int), cstream));
  B2_CUDA_OK(cudaMemcpyAsync(d_vals2, vals, (size_t)S.nnz * sizeof(double), cudaMemcpyHostToDevice, cstream));
  {
    const int nb = (int)std::min<int64_t>((S.nnz + 255) / 256, 2368);
    B2_LAUNCH(k_vals_compare, nb, 256, 0, cstream, S.nnz, S.nnz - S.nvar, d_vals, d_vals2, rho, d_mismatch);
  }
  B2_CUDA_OK(cudaMemcpyAsync(h_mismatch, d_mismatch, sizeof(int), cudaMemcpyDeviceToHost, cstream));
  B2_CUDA_OK(cudaEventRecord(ev_cmp, cstream));
  if (refactorize_shift(rho, std::nan(""), eig_tol, npos, nzero, nneg, breakdown)) return -1;
  B2_CUDA_OK(cudaEventSynchronize(ev_cmp));
  std::swap(d_vals, d_vals2);          // d_vals = what the caller holds now, in both cases
  if (*h_mismatch == 0) {
    if (speculation_held) *speculation_held = 1;
    return 0;
  }
  B2_CUDA_OK(cudaEventRecord(ev[0], stream));
  return assemble_and_factor(eig_tol, npos, nzero, nneg, breakdown, true);
}

int Engine::solve_core(const double* d_b, double* d_o, int negate, int refine_steps, double* relres) {
  const Symbolic& S = sym;
  const int64_t N = S.N;
  const int nb = (int)((N + 255) / 256);
  if (!factored) {
    snprintf(g_last_error, sizeof(g_last_error), "b2_solve before a factorization");
    return -2;
  }
  auto one_solve = [&](const double* b, int accumulate) -> int {
    B2_LAUNCH(k_perm_in, nb, 256, 0, stream, N, d_perm, b, d_x);
#ifndef B2_EMULATE
    if (use_graph) {
      if (!g_fwdbwd) {
        cudaGraph_t g = nullptr;
        B2_CUDA_OK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        int rc = run_solve_launches();
        cudaError_t ce = cudaStreamEndCapture(stream, &g);
        if (rc || ce != cudaSuccess) {
          if (!rc)
            snprintf(g_last_error, sizeof(g_last_error), "graph capture of the solve failed: %s", cudaGetErrorString(ce));
          if (ce == cudaSuccess && g) cudaGraphDestroy(g);
          return -1;
        }
        B2_CUDA_OK(cudaGraphInstantiate(&g_fwdbwd, g, 0));
        cudaGraphDestroy(g);
      }
      B2_CUDA_OK(cudaGraphLaunch(g_fwdbwd, stream));
    } else
#endif
    {
      if (run_solve_launches()) return -1;
    }
    B2_LAUNCH(k_perm_out, nb, 256, 0, stream, N, d_perm, d_x, d_sol, accumulate);
    return 0;
  };
  if (one_solve(d_b, 0)) return -1;
  last_sweeps = 1;
  const bool adaptive = refine_steps > 0 && refine_tol > 0;
  const bool need_res = refine_steps > 0 || relres != nullptr;
  auto norms = [&]() -> int {   // h_scalars[0] = ||res||^2, [1] = ||b||^2 (valid after a stream sync)
    int pb = (int)std::min<int64_t>(nb, 512);
    B2_LAUNCH(k_sumsq, pb, 256, 0, stream, N, d_res, d_part);
    B2_LAUNCH(k_fold, 1, 256, 0, stream, pb, d_part, d_part + 1024);
    B2_LAUNCH(k_sumsq, pb, 256, 0, stream, N, d_b, d_part);
    B2_LAUNCH(k_fold, 1, 256, 0, stream, pb, d_part, d_part + 1025);
    B2_CUDA_OK(cudaMemcpyAsync(h_scalars, d_part + 1024, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    return 0;
  };
  for (int it = 0; it <= refine_steps && need_res; it++) {
    B2_LAUNCH(k_residual, nb, 256, 0, stream, N, d_Sp, d_Sj, d_Sslot, d_nzval, d_sol, d_b, d_res);
    const bool last = it == refine_steps;
    if (adaptive && !last) {
      // refine only while the residual is above the tolerance (one 16-byte D2H + sync per check)
      if (norms()) return -1;
      B2_CUDA_OK(cudaStreamSynchronize(stream));
      const double rr = h_scalars[1] > 0 ? std::sqrt(h_scalars[0] / h_scalars[1]) : std::sqrt(h_scalars[0]);
      if (rr <= refine_tol) break;
    } else if (last) {
      if (relres && norms()) return -1;
      break;
    }
    if (one_solve(d_res, 1)) return -1;
    last_sweeps++;
  }
  B2_LAUNCH(k_scale_copy, nb, 256, 0, stream, N, d_sol, d_o, negate ? -1.0 : 1.0);
  B2_CUDA_OK(cudaGetLastError());
  return 0;
}

int Engine::solve_host(const double* rhs, double* out, int negate, int refine_steps, double* relres) {
  B2_CUDA_OK(cudaSetDevice(device));
  const size_t nbytes = (size_t)sym.N * sizeof(double);
  B2_CUDA_OK(cudaEventRecord(ev[0], stream));
  B2_CUDA_OK(cudaMemcpyAsync(d_rhs, rhs, nbytes, cudaMemcpyHostToDevice, stream));
  B2_CUDA_OK(cudaEventRecord(ev[1], stream));
  if (solve_core(d_rhs, d_out, negate, refine_steps, relres)) return -1;
  B2_CUDA_OK(cudaEventRecord(ev[4], stream));
  B2_CUDA_OK(cudaMemcpyAsync(out, d_out, nbytes, cudaMemcpyDeviceToHost, stream));
  B2_CUDA_OK(cudaEventRecord(ev[5], stream));
  B2_CUDA_OK(cudaStreamSynchronize(stream));
  if (relres) *relres = h_scalars[1] > 0 ? std::sqrt(h_scalars[0] / h_scalars[1]) : std::sqrt(h_scalars[0]);
  float ms = 0;
  cudaEventElapsedTime(&ms, ev[0], ev[1]); last_ms[0] = ms;
  cudaEventElapsedTime(&ms, ev[1], ev[4]); last_ms[3] = ms;
  cudaEventElapsedTime(&ms, ev[4], ev[5]); last_ms[4] = ms;
  return 0;
}

}  // namespace b2
