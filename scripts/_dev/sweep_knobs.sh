#!/bin/bash
# developer sweep: C4 factor / solve times over the engine's knobs (one at a time around the defaults)
out=gpurun_out/$1; mkdir -p $out
run() {
  tag=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu --no-batched --no-configs --preload 0 > $out/b_$tag.json 2> $out/b_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/b_$tag.json").read().strip().splitlines()[-1])
    print("%-28s value %.1f factor %.3f solve %.3f relres %.1e launches/step %d" % ("$tag", d["value"], d["phase_ms"]["factor"], d["phase_ms"]["solve"], d["relres"], d["gpu_launches"]/d["steps"]))
except Exception as e:
    print("$tag FAILED", e, open("$out/b_$tag.err").read()[-300:])
PY
}
run default X=1
run lm300 B2_DAG_LEVEL_MAX=300
run lm1000 B2_DAG_LEVEL_MAX=1000
run lm2000 B2_DAG_LEVEL_MAX=2000
run small56 B2_SMALL_MAX_M=56
run small96 B2_SMALL_MAX_M=96
run big64 B2_SOLVE_BIG_M=64
run big128 B2_SOLVE_BIG_M=128
run big192 B2_SOLVE_BIG_M=192
run inv1 B2_INV_MIN_BLK=1
run inv3 B2_INV_MIN_BLK=3
run fork1 B2_SOLVE_FORK=1
run relax15 B2_RELAX_SCALE=1.5
run relax07 B2_RELAX_SCALE=0.7
run tinys8 B2_TINY_SOLVE_MAX_M=8
run tinys32 B2_TINY_SOLVE_MAX_M=32
