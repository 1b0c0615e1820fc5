set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "nbody32 or ffnn" --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -25 gpurun_out/pytest_gpu.log
