# ncu --set full capture of the N-body kernel (one launch, 131072 lanes, propagate_until(20 yr)).
set -x
mkdir -p gpurun_out
for bt in 384 512; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_nb -c 1 -f -o gpurun_out/r2_k_nb_$bt python bench.py --no-cpu-baseline --batch 131072 --steps 1 --warmup 0 --tape nbody --lanes-per-thread 1 --block-threads $bt > gpurun_out/r2_k_nb_$bt.log 2>&1
grep -o '"lane_steps_per_step": [0-9]*' gpurun_out/r2_k_nb_$bt.log
done
ls -la gpurun_out/*.ncu-rep | tail -3
