set -x
mkdir -p gpurun_out
timeout 900 python tools/bench_configs.py > gpurun_out/r1_other_configs.jsonl 2> gpurun_out/r1_other_configs.err; cat gpurun_out/r1_other_configs.jsonl | cut -c1-400; tail -3 gpurun_out/r1_other_configs.err
