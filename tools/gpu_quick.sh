set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -9 gpurun_out/pytest_gpu.log
