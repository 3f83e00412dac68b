#!/bin/bash
TAG=${1:-r02ai}
OUT=gpurun_out/$TAG
mkdir -p $OUT
BNN_B200_TUNE_LOG=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dropin > $OUT/bench_r18.log 2> $OUT/tune_r18.log; echo "bench r18 $?"; tail -1 $OUT/bench_r18.log | cut -c1-200
grep -c "bnn tune" $OUT/tune_r18.log
