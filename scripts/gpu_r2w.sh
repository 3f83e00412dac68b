#!/bin/bash
# shortcut kernel phase-2 rework (register weights, carry-save triples, channel blocks over gridDim.y)
TAG=${1:-r02w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --config resnet50 --steps 20 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_r50.json > $OUT/bench_r50.log 2>&1; echo "bench r50 $?"; tail -1 $OUT/bench_r50.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_r18.json > $OUT/bench_r18.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18.log | cut -c1-200
python - <<'PY'
import json
for f in ("r50","r18"):
    d=json.load(open(f"gpurun_out/%s/layers_%s.json" % ("$TAG", f)))["per_layer"]
    print(f, {k: round(v["conv_ms"],4) for k,v in d.items() if "downsample" in k})
PY
