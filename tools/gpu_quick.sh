set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_random_expressions.py -m gpu -q > gpurun_out/pytest_random.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_random.log
