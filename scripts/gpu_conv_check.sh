#!/bin/bash
TAG=${1:-cc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest.log
for mode in mid out_cl; do timeout 300 python scripts/profile_layer.py --layers l1,l2,l3,l4 --reps 20 --fused $mode | tee -a $OUT/layers_iso.jsonl; done
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layers-out $OUT/layers.json > $OUT/bench.log 2>&1; tail -1 $OUT/bench.log | cut -c1-200
