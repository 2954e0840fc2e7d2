#!/bin/bash
# developer sweep: C4 factor/solve times for several extents of the dataflow launch
out=gpurun_out/$1; mkdir -p $out
for lm in 0 300 600 900 1500 2600 4000; do
  B2_DAG_LEVEL_MAX=$lm python bench.py --steps 10 --warmup 3 --no-cpu --no-batched > $out/bench_lm$lm.json 2> $out/bench_lm$lm.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/bench_lm$lm.json").read().strip().splitlines()[-1])
    print("level_max", $lm, "value", round(d["value"],1), d["phase_ms"], "relres", d["relres"], "launches/step", d["gpu_launches"]/d["steps"])
except Exception as e:
    print("level_max", $lm, "FAILED", e, open("$out/bench_lm$lm.err").read()[-400:])
PY
done
