#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
run() {
  tag=$1; shift
  env "$@" python bench.py --steps 10 --warmup 3 --no-cpu --no-batched --no-configs --preload 0 > $out/b_$tag.json 2> $out/b_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/b_$tag.json").read().strip().splitlines()[-1])
    print("%-36s value %.1f factor %.3f solve %.3f relres %.1e launches/step %d" % ("$tag", d["value"], d["phase_ms"]["factor"], d["phase_ms"]["solve"], d["relres"], d["gpu_launches"]/d["steps"]))
except Exception as e:
    print("$tag FAILED", e, open("$out/b_$tag.err").read()[-300:])
PY
}
run s96_fork_t8 B2_SMALL_MAX_M=96 B2_SOLVE_FORK=1 B2_TINY_SOLVE_MAX_M=8
run s112_fork_t8 B2_SMALL_MAX_M=112 B2_SOLVE_FORK=1 B2_TINY_SOLVE_MAX_M=8
run s128_fork_t8 B2_SMALL_MAX_M=128 B2_SOLVE_FORK=1 B2_TINY_SOLVE_MAX_M=8
run s96_fork_t4 B2_SMALL_MAX_M=96 B2_SOLVE_FORK=1 B2_TINY_SOLVE_MAX_M=4
run s96_fork_t12 B2_SMALL_MAX_M=96 B2_SOLVE_FORK=1 B2_TINY_SOLVE_MAX_M=12
run s96_fork_t8_r07 B2_SMALL_MAX_M=96 B2_SOLVE_FORK=1 B2_TINY_SOLVE_MAX_M=8 B2_RELAX_SCALE=0.7
run s96_fork_t8_r05 B2_SMALL_MAX_M=96 B2_SOLVE_FORK=1 B2_TINY_SOLVE_MAX_M=8 B2_RELAX_SCALE=0.5
run s96_fork_t8_big112 B2_SMALL_MAX_M=96 B2_SOLVE_FORK=1 B2_TINY_SOLVE_MAX_M=8 B2_SOLVE_BIG_M=112
run s96_fork_t8_big80 B2_SMALL_MAX_M=96 B2_SOLVE_FORK=1 B2_TINY_SOLVE_MAX_M=8 B2_SOLVE_BIG_M=80
run s96_fork_t8_lm1000 B2_SMALL_MAX_M=96 B2_SOLVE_FORK=1 B2_TINY_SOLVE_MAX_M=8 B2_DAG_LEVEL_MAX=1000
run s96_fork_t8_np3 B2_SMALL_MAX_M=96 B2_SOLVE_FORK=1 B2_TINY_SOLVE_MAX_M=8 B2_DAG_MIN_NP=3
run s96_fork_t8_tiny4 B2_SMALL_MAX_M=96 B2_SOLVE_FORK=1 B2_TINY_SOLVE_MAX_M=8 B2_TINY_MAX_M=4
