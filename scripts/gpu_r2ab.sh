#!/bin/bash
TAG=${1:-r02ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
BNN_B200_TUNE_LOG=1 timeout 600 python bench.py --config resnet50 --steps 10 --warmup 3 --no-cpu-baseline --no-dropin > $OUT/bench_r50.log 2> $OUT/tune_r50.log; echo "bench r50 $?"; tail -1 $OUT/bench_r50.log | cut -c1-200
grep -c "bnn tune" $OUT/tune_r50.log
