set -x
mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cut -c1-200 gpurun_out/bench_a.json
timeout 600 python -m pytest tests -m gpu -x -q -k "outer_ss or closed_form" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
