set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cpp-e2e --batch 262144 2>> gpurun_out/r2_iter.err | tee gpurun_out/r2_iter.jsonl | cut -c1-120
timeout 300 python tools/bench_configs.py tb 2>> gpurun_out/r2_iter.err | tee -a gpurun_out/r2_iter.jsonl | cut -c1-100,200-330
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nbody" > gpurun_out/r2_pytest_iter.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_iter.log | cut -c1-300
