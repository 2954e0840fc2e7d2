mkdir -p gpurun_out/r2n; O=gpurun_out/r2n
S=$(date +%s); python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? wall $(( $(date +%s) - S )) s"; tail -3 $O/pytest_gpu.log
python scripts/profile_launches.py c4 512 0 v > $O/warm_c4.txt 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-configs --preload 0 > $O/bench_under_ncu.log 2>&1
echo "traffic rc=$?"; wc -l $O/traffic.csv
ncu --set full --clock-control none --import-source on -k regex:k_nls_dense -s 1 -c 1 -o $O/prof_nls python scripts/_dev/nls_ncu.py > $O/nls_ncu.log 2>&1; echo "nls ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_front_dag -s 3 -c 1 -o $O/prof_dag python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-configs --preload 0 > $O/dag_ncu.log 2>&1; echo "dag ncu rc=$?"
ls -la $O
