#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
run() {
  tag=$1; shift
  env "$@" python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu --no-batched --no-configs --preload 0 > $out/b_$tag.json 2> $out/b_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/b_$tag.json").read().strip().splitlines()[-1])
    print("%-24s value %.1f factor %.3f solve %.3f relres %.1e sweeps %d" % ("$tag", d["value"], d["phase_ms"]["factor"], d["phase_ms"]["solve"], d["relres"], d["config"]["solve_sweeps_used"]))
except Exception as e:
    print("$tag FAILED", e, open("$out/b_$tag.err").read()[-300:])
PY
}
run alldag B2_DAG_LEVEL_MAX=1000000000 B2_DAG_MIN_NP=1
run lm20000_np1 B2_DAG_LEVEL_MAX=20000 B2_DAG_MIN_NP=1
run lm100000_np1 B2_DAG_LEVEL_MAX=100000 B2_DAG_MIN_NP=1
