# Round 2: the dense-network kernel: parity tests, timing, one ncu capture.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "ffnn" > gpurun_out/r2_pytest_nn.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_nn.log | cut -c1-300
timeout 600 python tools/bench_configs.py nn > gpurun_out/r2_nn_configs.jsonl 2> gpurun_out/r2_nn_configs.err; cut -c1-330 gpurun_out/r2_nn_configs.jsonl; tail -3 gpurun_out/r2_nn_configs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nn -c 1 -o gpurun_out/r2_k_nn -f python tools/bench_configs.py nnsmall > gpurun_out/r2_k_nn.log 2>&1; tail -2 gpurun_out/r2_k_nn.log | cut -c1-300
