#!/bin/bash
TAG=${1:-fc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest.log
timeout 600 python scripts/bench_configs.py ${CFGS:-resnet50_xnorpp} 2>&1 | tail -2 | tee $OUT/configs.jsonl
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > $OUT/bench.log 2>&1; tail -1 $OUT/bench.log | cut -c1-200
