#!/bin/bash
TAG=${1:-pack}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest.log
timeout 600 python scripts/bench_configs.py hblock_net 2>&1 | tail -1 | tee $OUT/configs.jsonl
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-fuse > $OUT/bench_nofuse.log 2>&1; tail -1 $OUT/bench_nofuse.log | cut -c1-200
