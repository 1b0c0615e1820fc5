set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_events.py -m gpu -q --maxfail=30 > gpurun_out/r2_pytest_ev.log 2>&1; echo "pytest ev rc=$?"; tail -40 gpurun_out/r2_pytest_ev.log
timeout 600 python -m pytest tests/test_cpp_api.py -m gpu -q > gpurun_out/r2_pytest_cpp.log 2>&1; echo "pytest cpp rc=$?"; grep -E "d_output|REQUIRE|passed|failed" gpurun_out/r2_pytest_cpp.log | head -20
