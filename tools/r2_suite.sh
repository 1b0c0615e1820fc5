# Full GPU suite + smoke.
set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
timeout 2400 python -m pytest tests -m gpu -q --durations=5 --maxfail=8 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -10 gpurun_out/r2_pytest_gpu.log
