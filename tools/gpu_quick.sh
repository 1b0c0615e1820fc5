set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "kepler" --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
