"""Developer aid: timeline of the tasks of ONE front of the k_front_dag launch (after its last diagonal block)."""
import os, sys
sys.argv = [sys.argv[0]] + sys.argv[1:]
lvl = int(sys.argv[2]) if len(sys.argv) > 2 else 13
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dag_trace2.py")).read().split("# the widest front of every level")[0])
q = L == lvl
fs = np.unique(front[q])
f = fs[np.argmax(w[fs])]
sel = front == f
a = tr[sel]; ii = I[sel]; jj = J[sel]; ch = chain[sel]
npb = (w[f] + 63) // 64
print(f"front {f}: m {m[f]} w {w[f]} np {npb}; tasks {sel.sum()}")
order = np.argsort(a[:, 9])
print(" I  J kind  sm   start  chdone  asmbld  upd_done  subst_wait subst_ld subst_done ldlt0 ldlt1   end   dur")
for k in order:
    d = a[k]
    g = lambda x: f"{us(x):8.1f}" if x > 0 else "       -"
    print(f"{ii[k]:2d} {jj[k]:2d} {'chain' if ch[k] else '     '} {d[1]:3d} {g(d[2])} {g(d[10])} {g(d[11])} {g(d[3])} {g(d[4])} {g(d[5])} {g(d[6])} {g(d[7])} {g(d[8])} {g(d[9])} {(d[9]-d[2])/1e3:6.1f}")
# parent's first task
