#!/bin/bash
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_kernels.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_kernel -s 1 -c 1 \
    -o $OUT/prof_stem -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $OUT/ncu_stem.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bconv_kernel -s 1 -c 1 \
    -o $OUT/prof_l1_out -f python scripts/profile_layer.py --layers l1 --reps 2 --fused out_cl > $OUT/ncu_l1o.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 60 --csv \
    --log-file $OUT/launches_head.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $OUT/bench.log 2>&1; tail -1 $OUT/bench.log | cut -c1-200
ls $OUT
