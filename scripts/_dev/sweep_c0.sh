#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
run() {
  wl=$1; tag=$2; shift; shift
  env "$@" python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu --no-batched --no-configs --preload 0 > $out/b_${wl}_$tag.json 2> $out/b_${wl}_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("$out/b_${wl}_$tag.json").read().strip().splitlines()[-1])
    print("%-4s %-10s value %.1f factor %.3f solve %.3f relres %.1e" % ("$wl", "$tag", d["value"], d["phase_ms"]["factor"], d["phase_ms"]["solve"], d["relres"]))
except Exception as e:
    print("$wl $tag FAILED", e, open("$out/b_${wl}_$tag.err").read()[-300:])
PY
}
for v in 16 24 32; do run c4 c0_$v B2_SMALL_C0_MAX=$v; done
for v in 16 24 32; do run c3 c0_$v B2_SMALL_C0_MAX=$v; done
run c2 c0_16 B2_SMALL_C0_MAX=16
run c2 c0_24 B2_SMALL_C0_MAX=24
