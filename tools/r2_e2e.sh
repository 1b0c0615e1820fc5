set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_pytest_e2e.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_e2e.log | cut -c1-300
rm -f gpurun_out/r2_bench_e2e.jsonl
for sub in 1 2 4 8; do
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cpp-e2e --e2e-sub $sub 2>> gpurun_out/r2_bench_e2e.err | tee -a gpurun_out/r2_bench_e2e.jsonl | cut -c1-100
done
