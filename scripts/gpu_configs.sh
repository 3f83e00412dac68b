#!/bin/bash
TAG=${1:-cfgs}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python scripts/bench_configs.py 2>&1 | tail -6 | tee $OUT/configs.jsonl
