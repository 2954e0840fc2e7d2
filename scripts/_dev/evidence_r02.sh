# round-2 evidence: GPU tests, launch list with DRAM traffic, ncu --set full of the top kernels, bench lines
mkdir -p gpurun_out/r3t; O=gpurun_out/r3t
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python scripts/profile_launches.py c4 512 0 v > $O/warm_c4.txt 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-configs --preload 0 > $O/bench_under_ncu.log 2>&1; echo "traffic rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_nls_dense -s 1 -c 1 -o $O/prof_nls python scripts/_dev/nls_ncu.py > $O/nls_ncu.log 2>&1; echo "nls ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_batched -s 2 -c 1 -o $O/prof_batched python scripts/_dev/batched_ncu.py > $O/batched_ncu.log 2>&1; echo "batched ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_front_dag -s 3 -c 1 -o $O/prof_dag python bench.py --steps 1 --warmup 1 --no-cpu --no-batched --no-configs --preload 0 > $O/dag_ncu.log 2>&1; echo "dag ncu rc=$?"
python scripts/_dev/batched_timing.py > $O/batched_timing.txt 2>&1
ls -la $O
