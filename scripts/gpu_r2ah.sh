#!/bin/bash
# conversion-free dot -> float in every epilogue form (EPI 0 / 1 / 2 too): parity suite, HBlock and ResNet-18 (dropin) benches
TAG=${1:-r02ah}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --config hblock --steps 30 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_hb.json > $OUT/bench_hb.log 2>&1; echo "bench hb $?"; tail -1 $OUT/bench_hb.log | cut -c1-200
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_r18.json > $OUT/bench_r18.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18.log | cut -c1-200
python - <<PY
import json
for c in ("hb","r18"):
    d=json.load(open("$OUT/layers_%s.json" % c))["line"]
    print(c, round(d["ms_per_step"],4), "dropin", round(d["dropin"]["ms_per_step"],3), round(d["dropin"]["value"],1), "roofline", round(d["roofline"]["achieved"],1))
PY
