# Round-2 round-end sequence: smoke, bench (reference arm + ours), other configurations, ncu launch list, sanitizer on
# the new kernels, full GPU suite.
set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_final.err; cut -c1-300 gpurun_out/r2_bench_reference.json
timeout 900 python bench.py > gpurun_out/r2_bench.json 2>> gpurun_out/r2_bench_final.err; cut -c1-300 gpurun_out/r2_bench.json
timeout 900 python tools/bench_configs.py tb n32 nn grid s6long > gpurun_out/r2_other_configs.jsonl 2> gpurun_out/r2_other_configs.err; cut -c1-100,200-360 gpurun_out/r2_other_configs.jsonl; tail -3 gpurun_out/r2_other_configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cpp-e2e > gpurun_out/r2_launches_bench.log 2>&1
tail -3 gpurun_out/r2_launches.csv | cut -c1-300
tail -3 gpurun_out/r2_bench_final.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_body_lane_kernel or (ffnn_parity and nn) or (step_parity_two_body and nbody-lane)" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_memcheck.log | cut -c1-200
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(ffnn_parity and nn) or (step_parity_two_body and nbody-lane)" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_sanitizer_racecheck.log | cut -c1-200
timeout 2400 python -m pytest tests -m gpu -q --durations=8 --maxfail=8 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_gpu.log
