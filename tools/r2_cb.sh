# Round 2: step callbacks in propagate_grid() and together with continuous output (C++ class + Python front end).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cpp_api.py tests/test_gpu_multi.py tests/test_gpu_events.py -m gpu -q -x -k "test_gpu_events or test_gpu_multi or cpp_api" > gpurun_out/r2_pytest_cb.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_cb.log | cut -c1-400
