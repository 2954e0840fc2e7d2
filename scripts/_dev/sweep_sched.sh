#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
run() {
  tag=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu --no-batched --no-configs --preload 0 > $out/b_$tag.json 2> $out/b_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/b_$tag.json").read().strip().splitlines()[-1])
    print("%-30s value %.1f factor %.3f solve %.3f relres %.1e" % ("$tag", d["value"], d["phase_ms"]["factor"], d["phase_ms"]["solve"], d["relres"]))
except Exception as e:
    print("$tag FAILED", e, open("$out/b_$tag.err").read()[-300:])
PY
}
run default X=1
run tupd4 B2_DAG_TUPD=4
run tupd5.5 B2_DAG_TUPD=5.5
run tupd7 B2_DAG_TUPD=7
run tupd10 B2_DAG_TUPD=10
run tasm4 B2_DAG_TASM=4
run tasm12 B2_DAG_TASM=12
run tupd5.5_tasm4 B2_DAG_TUPD=5.5 B2_DAG_TASM=4
run tupd7_tasm5 B2_DAG_TUPD=7 B2_DAG_TASM=5
run tupd1 B2_DAG_TUPD=1
run nosched B2_DAG_SCHED=0
