set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "nbody32 or ffnn or global or kernel_selection" --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -25 gpurun_out/pytest_gpu.log
