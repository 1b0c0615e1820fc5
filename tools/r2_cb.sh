# Round 2: API additions (callbacks / continuous output / events / shards): their tests.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_pytest_cb.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_cb.log | cut -c1-600
