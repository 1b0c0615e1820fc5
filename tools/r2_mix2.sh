set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_events.py tests/test_cpp_api.py tests/test_nb_plan.py -m gpu -q --maxfail=30 > gpurun_out/r2_pytest_ev.log 2>&1; echo "pytest ev rc=$?"; tail -8 gpurun_out/r2_pytest_ev.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lean_division or nbody32 or two_body or outer_ss" > gpurun_out/r2_pytest_nbdiv.log 2>&1; echo "pytest nbdiv rc=$?"; tail -5 gpurun_out/r2_pytest_nbdiv.log | cut -c1-300
timeout 600 python tools/bench_configs.py tb n32 > gpurun_out/r2_other_configs_v4.jsonl 2> gpurun_out/r2_other_configs_v4.err; cut -c1-330 gpurun_out/r2_other_configs_v4.jsonl
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cpp-e2e > gpurun_out/r2_bench_div.jsonl 2> gpurun_out/r2_bench_div.err; cut -c1-200 gpurun_out/r2_bench_div.jsonl
