import sys,os,time,functools,subprocess
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT)
if len(sys.argv) < 3:
    for ng in ("0", "1"):
        os.environ["B2_NO_GRAPH"] = ng
        print("B2_NO_GRAPH", ng, flush=True)
        for big in ("32",):
            for size in ("64","160"):
                for order in ("3","0"):
                    subprocess.run([sys.executable, __file__, big, size, order])
    sys.exit(0)
os.environ["B2_SOLVE_BIG_M"]=sys.argv[1]
import numpy as np
from cannoles_b200.linsolve import B200Struct
from cannoles_b200.workloads import first_system, make_config
EPS=2.0**-52
order=int(sys.argv[3])
nls,method,desc=make_config("c4",int(sys.argv[2]))
s,rhs=first_system(nls,method,functools.partial(B200Struct,nvar=nls.nvar,nequ=nls.nequ,ncon=nls.ncon,ordering=order,refine_steps=0,shift_retries=False))
B=s.LDLT
ok=B.try_to_factorize(s.vals,nls.nvar,nls.nequ,nls.ncon,EPS)
d=np.zeros(B.N)
res=[]
for rep in range(4):
    B.solve_ldl(rhs * (1 + rep) + rep * np.roll(rhs, rep),d); res.append(B.last_relres)
st=B.stats()
print("big_m",sys.argv[1],"size",sys.argv[2],"order",order,"ok",ok,"max_front",st["max_front"],"relres",res, "nan" if np.isnan(d).any() else "")
