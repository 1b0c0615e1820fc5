set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cpp_api.py -m gpu -q -x -k "grid or cpp_api" > gpurun_out/r2_pytest_grid.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_grid.log | cut -c1-300
timeout 600 python tools/bench_configs.py grid > gpurun_out/r2_grid.jsonl 2> gpurun_out/r2_grid.err; cut -c1-400 gpurun_out/r2_grid.jsonl; tail -3 gpurun_out/r2_grid.err
