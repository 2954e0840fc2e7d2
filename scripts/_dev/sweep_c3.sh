#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
run() {
  tag=$1; shift
  env "$@" python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu --no-batched --no-configs --preload 0 > $out/b_$tag.json 2> $out/b_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/b_$tag.json").read().strip().splitlines()[-1])
    print("%-24s value %.1f factor %.3f solve %.3f relres %.1e sweeps %d levels %d max_front %d" % ("$tag", d["value"], d["phase_ms"]["factor"], d["phase_ms"]["solve"], d["relres"], d["config"]["solve_sweeps_used"], d["config"]["nlevels"], d["config"]["max_front"]))
except Exception as e:
    print("$tag FAILED", e, open("$out/b_$tag.err").read()[-300:])
PY
}
run default X=1
run np2 B2_DAG_MIN_NP=2
run np3 B2_DAG_MIN_NP=3
run lm3000 B2_DAG_LEVEL_MAX=3000
run lm3000_np2 B2_DAG_LEVEL_MAX=3000 B2_DAG_MIN_NP=2
run big128 B2_SOLVE_BIG_M=128
run big160 B2_SOLVE_BIG_M=160
run inv1 B2_INV_MIN_BLK=1
run nofork B2_SOLVE_FORK=-1
