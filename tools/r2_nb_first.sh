# Round 2, first GPU check of the N-body kernel: smoke, the N-body-shaped parity tests, short bench lines per shape.
set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "nbody or outer_ss or two_body or kernel_selection" > gpurun_out/r2_pytest_nb.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_nb.log
for cfg in "--tape auto" "--tape nbody --lanes-per-thread 1 --block-threads 384" "--tape nbody --lanes-per-thread 2" "--tape smem"; do
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch 262144 $cfg 2>> gpurun_out/r2_bench_first.err | tee -a gpurun_out/r2_bench_first.jsonl | cut -c1-330
done
