# ncu --set full capture of the N-body kernel with CTA teams (model::nbody N = 32, 8192 lanes, propagate_until(1)).
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nb -c 1 -f -o gpurun_out/r2_k_nb_cta_n32 python tools/bench_configs.py n32 > gpurun_out/r2_k_nb_cta_n32.log 2>&1; tail -3 gpurun_out/r2_k_nb_cta_n32.log | cut -c1-300
