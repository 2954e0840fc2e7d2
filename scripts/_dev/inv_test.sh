mkdir -p gpurun_out/r2r; O=gpurun_out/r2r
for v in 0 2 3 4; do
B2_INV_MIN_BLK=$v python bench.py --steps 20 --warmup 3 --no-cpu --no-batched --no-configs --preload 0.5 > $O/bench_inv$v.json 2> $O/err$v.txt
python - <<PY
import json
d=json.loads(open("$O/bench_inv$v.json").read().strip().splitlines()[-1])
print("inv_min_blk", $v, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["phase_ms"], "relres", d["relres"], "sweeps", d["config"]["solve_sweeps_used"])
PY
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/profile_launches.py c4 512 0 v > $O/warm_c4.txt 2>&1
python scripts/_dev/batched_timing.py > $O/batched_timing.txt 2>&1; cat $O/batched_timing.txt
