set -x
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cut -c1-200 gpurun_out/bench_a.json
HEYOKA_B200_TMEM_ROWS=2 timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_b.json 2>> gpurun_out/bench_a.err; cut -c1-200 gpurun_out/bench_b.json
timeout 300 python bench.py --no-cpu-baseline --lanes-per-warp 2 --lanes-per-thread 1 --block-threads 448 --steps 3 > gpurun_out/bench_c.json 2>> gpurun_out/bench_a.err; cut -c1-200 gpurun_out/bench_c.json
timeout 300 python bench.py --no-cpu-baseline --lanes-per-warp 1 --lanes-per-thread 1 --steps 3 > gpurun_out/bench_d.json 2>> gpurun_out/bench_a.err; cut -c1-200 gpurun_out/bench_d.json
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_coop -c 1 -o gpurun_out/prof_r1_coop_v8 -f python bench.py --no-cpu-baseline --batch 131072 --steps 1 --warmup 0 > gpurun_out/ncu_v8.log 2>&1
tail -2 gpurun_out/ncu_v8.log | cut -c1-300
tail -5 gpurun_out/bench_a.err
