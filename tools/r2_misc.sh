set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "reference_default_mode" > gpurun_out/r2_pytest_misc.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_misc.log | cut -c1-300
timeout 600 python tools/bench_configs.py s6long > gpurun_out/r2_s6long.jsonl 2> gpurun_out/r2_s6long.err; cut -c1-100,200-400 gpurun_out/r2_s6long.jsonl; tail -2 gpurun_out/r2_s6long.err
