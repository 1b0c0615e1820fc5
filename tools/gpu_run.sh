set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "continuous_output or cpp_api or grid" --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -25 gpurun_out/pytest_gpu.log
