# Round 2: full GPU test suite + default bench + the other configs.
set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
timeout 2400 python -m pytest tests -m gpu -q --durations=8 --maxfail=8 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest_gpu.log
rm -f gpurun_out/r2_bench_third.jsonl
for cfg in "--tape auto" "--tape nbody --lanes-per-thread 1 --block-threads 512"; do
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch 262144 $cfg 2>> gpurun_out/r2_bench_third.err | tee -a gpurun_out/r2_bench_third.jsonl | cut -c1-130
done
timeout 600 python tools/bench_configs.py tb n32 > gpurun_out/r2_other_configs_v3.jsonl 2> gpurun_out/r2_other_configs_v3.err; cut -c1-330 gpurun_out/r2_other_configs_v3.jsonl
