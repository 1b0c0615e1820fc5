set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r1_bench_2gpu.json 2> gpurun_out/r1_bench_2gpu.err; echo rc=$?; cut -c1-400 gpurun_out/r1_bench_2gpu.json; tail -5 gpurun_out/r1_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r1_bench_2gpu_ref.json 2>> gpurun_out/r1_bench_2gpu.err; echo rc=$?; cut -c1-300 gpurun_out/r1_bench_2gpu_ref.json
