# Round 2: N-body kernel v2 (roles, n-slots, norms in shared memory): parity subset, bench lines per shape, other configs, ncu.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "nbody or outer_ss or two_body or kernel_selection or dense_output or global_exits" > gpurun_out/r2_pytest_nb2.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_nb2.log
rm -f gpurun_out/r2_bench_second.jsonl
for cfg in "--tape auto" "--tape nbody --lanes-per-thread 1 --block-threads 384" "--tape nbody --lanes-per-thread 1 --block-threads 256" "--tape nbody --lanes-per-thread 2"; do
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch 262144 $cfg 2>> gpurun_out/r2_bench_second.err | tee -a gpurun_out/r2_bench_second.jsonl | cut -c1-130
done
timeout 600 python tools/bench_configs.py tb n32 > gpurun_out/r2_other_configs_v2.jsonl 2> gpurun_out/r2_other_configs_v2.err; cut -c1-400 gpurun_out/r2_other_configs_v2.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nb -c 1 -f -o gpurun_out/r2_k_nb_v2_384 python bench.py --no-cpu-baseline --batch 131072 --steps 1 --warmup 0 --tape nbody --lanes-per-thread 1 --block-threads 384 > gpurun_out/r2_k_nb_v2_384.log 2>&1
