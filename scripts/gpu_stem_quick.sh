#!/bin/bash
TAG=${1:-stemq}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python scripts/time_stem.py 2>&1 | tail -3 | tee $OUT/time_stem.json
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "stem" 2>&1 | tail -2
