set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cpp_api.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
