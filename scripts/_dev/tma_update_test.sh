mkdir -p gpurun_out/r4h; O=gpurun_out/r4h
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
B2_DAG=0 python -m pytest tests/test_gpu_parity.py -x -q -k "slice or full_size or known" 2>&1 | tail -2
for wl in c4 c3; do python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-batched --no-configs --preload 0 2>$O/err_$wl.txt | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('$wl value %.1f factor %.3f solve %.3f relres %.1e' % (d['value'], d['phase_ms']['factor'], d['phase_ms']['solve'], d['relres']))
"; done
B2_DAG=0 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-batched --no-configs --preload 0 2>>$O/err_c4.txt | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('c4 DAG=0 value %.1f factor %.3f solve %.3f relres %.1e' % (d['value'], d['phase_ms']['factor'], d['phase_ms']['solve'], d['relres']))
"
B2_DAG=0 python scripts/_dev/determinism_loop.py c3 5000 300 | tail -1
python scripts/profile_launches.py c3 50000 0 2>&1 | grep -E "update|trsm|assemble_large" | head -5
