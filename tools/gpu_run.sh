# Scratch sequence for a gpurun call: smoke, the GPU test suite, one short bench line.
set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cut -c1-200 gpurun_out/bench_a.json
