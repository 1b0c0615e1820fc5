set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python tools/bench_multi.py n32 s6strong > gpurun_out/r2_multi_8gpu.jsonl 2> gpurun_out/r2_multi_8gpu.err; cut -c1-400 gpurun_out/r2_multi_8gpu.jsonl; tail -3 gpurun_out/r2_multi_8gpu.err
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2_pytest_8gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_8gpu.log | cut -c1-300
