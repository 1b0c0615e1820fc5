# Round 2, two GPUs: sharded-batch tests, bench.py under torchrun (both arms), grid timing.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_cpp_api.py -m gpu -q > gpurun_out/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_2gpu.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo rc=$?; cut -c1-300 gpurun_out/r2_bench_2gpu.json; tail -3 gpurun_out/r2_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_2gpu_ref.json 2>> gpurun_out/r2_bench_2gpu.err; echo rc=$?; cut -c1-300 gpurun_out/r2_bench_2gpu_ref.json
timeout 600 python tools/bench_configs.py grid > gpurun_out/r2_grid.jsonl 2> gpurun_out/r2_grid.err; cut -c1-400 gpurun_out/r2_grid.jsonl; tail -3 gpurun_out/r2_grid.err
