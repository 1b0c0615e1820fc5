# Round 2: the dense-network kernel: parity, timing, ncu capture; then the full GPU suite.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ffnn" > gpurun_out/r2_pytest_nn.log 2>&1; echo "pytest nn rc=$?"; tail -15 gpurun_out/r2_pytest_nn.log
timeout 600 python tools/bench_configs.py nn > gpurun_out/r2_nn_configs.jsonl 2> gpurun_out/r2_nn_configs.err; cut -c1-330 gpurun_out/r2_nn_configs.jsonl; tail -5 gpurun_out/r2_nn_configs.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nn -c 1 -o gpurun_out/r2_k_nn -f python tools/bench_configs.py nnsmall > gpurun_out/r2_k_nn.log 2>&1; tail -3 gpurun_out/r2_k_nn.log
timeout 2400 python -m pytest tests -m gpu -q --durations=8 --maxfail=8 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest_gpu.log
