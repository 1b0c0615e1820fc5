set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cut -c1-200 gpurun_out/bench_a.json
HEYOKA_B200_FUSE_SV=0 python bench.py --no-cpu-baseline > gpurun_out/bench_b.json 2>> gpurun_out/bench_a.err; cut -c1-200 gpurun_out/bench_b.json
python bench.py --no-cpu-baseline --lanes-per-warp 2 --lanes-per-thread 1 --steps 2 > gpurun_out/bench_c.json 2>> gpurun_out/bench_a.err; cut -c1-200 gpurun_out/bench_c.json
ncu --set full --clock-control none --import-source on -k regex:k_coop -c 1 -o gpurun_out/prof_r1_coop_v6 -f python bench.py --no-cpu-baseline --batch 131072 --steps 1 --warmup 0 > gpurun_out/ncu_v6.log 2>&1
tail -2 gpurun_out/ncu_v6.log | cut -c1-300
