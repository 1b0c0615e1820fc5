# Round-end measurement sequence: bench (ours + reference arm), ncu launch list, one full ncu capture.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; cut -c1-300 gpurun_out/r1_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench.err; cut -c1-300 gpurun_out/r1_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_launches_bench.log 2>&1
tail -3 gpurun_out/r1_launches.csv | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_coop -c 1 -o gpurun_out/r1_k_coop_full -f python bench.py --no-cpu-baseline --batch 131072 --steps 1 --warmup 0 > gpurun_out/r1_k_coop_full.log 2>&1
grep -o '"lane_steps_per_step": [0-9]*' gpurun_out/r1_k_coop_full.log
tail -3 gpurun_out/r1_bench.err
