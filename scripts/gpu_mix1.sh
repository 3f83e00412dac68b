#!/bin/bash
TAG=${1:-mix}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "pipeline or fused_engine" 2>&1 | tail -2
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench.log 2>&1; python - <<PY
import json
l=json.loads(open("$OUT/bench.log").read().strip().splitlines()[-1])
print("value",l["value"],"ms",l["ms_per_step"],"e2e",l["e2e"]["value"],l["e2e"]["ms_per_step"])
PY
LAYER=l1 MODES=out_cl bash scripts/gpu_prof_conv.sh $TAG | tail -2
