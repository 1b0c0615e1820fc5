# ncu --set full captures of the current kernels: S6 (k_nb, 131072 lanes, propagate_until(20 yr)), two-body (k_nb1, 2^22 lanes).
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nb -c 1 -f -o gpurun_out/r2_k_nb_v3_384 python bench.py --no-cpu-baseline --no-cpp-e2e --batch 131072 --steps 1 --warmup 0 > gpurun_out/r2_k_nb_v3_384.log 2>&1
grep -o '"lane_steps_per_step": [0-9]*' gpurun_out/r2_k_nb_v3_384.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nb1 -c 1 -o gpurun_out/r2_k_nb1_tb -f python tools/bench_configs.py tbsmall > gpurun_out/r2_k_nb1_tb.log 2>&1; tail -2 gpurun_out/r2_k_nb1_tb.log | cut -c1-300
