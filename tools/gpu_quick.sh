set -x
mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cut -c1-160 gpurun_out/bench_a.json
timeout 900 python -m pytest tests -m gpu -x -q -k "outer_ss or two_body or kepler or nbody32 or global_exits or kernel_selection" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
