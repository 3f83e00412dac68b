#!/bin/bash
TAG=${1:-sc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python scripts/time_shortcut.py 2>&1 | tail -4 | tee $OUT/time_shortcut.jsonl
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "shortcut" 2>&1 | tail -2
