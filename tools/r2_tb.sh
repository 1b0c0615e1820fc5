# Round 2: the one-thread-per-lane N-body kernel (k_nb1): parity tests, two-body timings, one ncu capture.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "two_body or two_massive or lean_division or kernel_selection" > gpurun_out/r2_pytest_tb.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest_tb.log | cut -c1-300
timeout 600 python tools/bench_configs.py tb tbteam tbsweep > gpurun_out/r2_tb.jsonl 2> gpurun_out/r2_tb.err; cut -c1-330 gpurun_out/r2_tb.jsonl; tail -3 gpurun_out/r2_tb.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nb1 -c 1 -o gpurun_out/r2_k_nb1_tb -f python tools/bench_configs.py tbsmall > gpurun_out/r2_k_nb1_tb.log 2>&1; tail -2 gpurun_out/r2_k_nb1_tb.log | cut -c1-300
