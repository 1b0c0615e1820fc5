set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -k "nbody32 or ffnn or global or kernel_selection" --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -14 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_configs.py n32 nn > gpurun_out/other2.jsonl 2> gpurun_out/other2.err; cut -c1-140 gpurun_out/other2.jsonl; python - <<'PY'
import json
for l in open('gpurun_out/other2.jsonl'):
    d=json.loads(l); print(d['config'][:70], d['kernel'], round(d['seconds'],4), int(d['lane_steps_per_s']), round(d['b_tape_gbs'],1))
PY
tail -3 gpurun_out/other2.err
timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cut -c1-200 gpurun_out/bench_a.json
