#!/bin/bash
# Host-side C++ of the library (expression system, decomposition, lowering, planners, C ABI, the drop-in class) rebuilt
# with AddressSanitizer + UndefinedBehaviorSanitizer and linked with the already-built CUDA objects; the CPU test suite
# then runs against that library. No GPU needed. Usage: bash tools/host_sanitize.sh  (from the repository root, after
# `python -m heyoka_b200.build`). Restores the real library afterwards; the log goes to stdout.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
cd "$ROOT/heyoka_b200/csrc"
for f in expression decompose model lower smem_plan nb_plan nn_plan capi_host taylor_adaptive_batch; do
    g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -I../../include -I. \
        -I/usr/local/cuda/include -c $f.cpp -o $W/$f.o &
done
wait
cd "$ROOT"
g++ -shared -fsanitize=address,undefined -o $W/libheyoka_b200.so $W/*.o $(ls build/obj/*.o | grep -v "\.cpp\.o") \
    -L/usr/local/cuda/lib64 -lcudart
LIB=heyoka_b200/lib/libheyoka_b200.so
cp -p $LIB $W/real.so
trap 'cp -p $W/real.so $LIB; rm -rf $W' EXIT
cp $W/libheyoka_b200.so $LIB
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) \
ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:log_path=$W/asan.log UBSAN_OPTIONS=print_stacktrace=1:log_path=$W/ubsan.log \
python -m pytest tests/test_host.py tests/test_oracle_golden.py tests/test_random_expressions.py tests/test_events_cpu.py \
    tests/test_codegen_cpu.py tests/test_nb_plan.py tests/test_multiprocess_cpu.py -q -m "not gpu" 2>&1 | tail -3
n=$(cat $W/asan.log* $W/ubsan.log* 2>/dev/null | grep -c "runtime error\|AddressSanitizer" || true)
echo "sanitizer reports: $n"
cat $W/asan.log* $W/ubsan.log* 2>/dev/null | head -50
