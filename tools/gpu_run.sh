set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 300 python tools/bench_configs.py tb > gpurun_out/tb.jsonl 2> gpurun_out/tb.err; cut -c1-330 gpurun_out/tb.jsonl; tail -2 gpurun_out/tb.err
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -16 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cut -c1-200 gpurun_out/bench_a.json
