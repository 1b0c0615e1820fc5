set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_cpp_api.py -m gpu -q > gpurun_out/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_2gpu.log | cut -c1-300
timeout 600 python tools/bench_multi.py n32 check s6strong > gpurun_out/r2_multi_2gpu.jsonl 2> gpurun_out/r2_multi_2gpu.err; cut -c1-400 gpurun_out/r2_multi_2gpu.jsonl; tail -3 gpurun_out/r2_multi_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo rc=$?; cut -c1-300 gpurun_out/r2_bench_2gpu.json; tail -3 gpurun_out/r2_bench_2gpu.err
