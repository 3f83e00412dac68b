#!/bin/bash
TAG=${1:-stem}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python scripts/time_stem.py 2>&1 | tail -3 | tee $OUT/time_stem.json
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -s -k "stem or fused_engine" > $OUT/pytest_stem.log 2>&1; echo "pytest exit $?"; grep -E "stem|passed|failed|rror" $OUT/pytest_stem.log | tail -25
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layers-out $OUT/layers.json > $OUT/bench.log 2>&1; tail -1 $OUT/bench.log | cut -c1-260
