#!/bin/bash
TAG=${1:-sc2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python scripts/bench_configs.py resnet50_xnorpp 2>&1 | tail -1 | tee $OUT/configs.jsonl
