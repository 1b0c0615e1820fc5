set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_events.py tests/test_cpp_api.py -m gpu -q --maxfail=30 > gpurun_out/r2_pytest_ev.log 2>&1; echo "pytest ev rc=$?"; tail -8 gpurun_out/r2_pytest_ev.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_cpp.jsonl 2> gpurun_out/r2_bench_cpp.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_cpp.jsonl').read().splitlines()[-1]); print(d['value'], d['e2e'], d.get('e2e_cpp_class'))"; tail -3 gpurun_out/r2_bench_cpp.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nb -c 1 -o gpurun_out/r2_k_nb_tb -f python tools/bench_configs.py tbsmall > gpurun_out/r2_k_nb_tb.log 2>&1; tail -3 gpurun_out/r2_k_nb_tb.log | cut -c1-300
