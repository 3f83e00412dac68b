#!/bin/bash
# HBlock engine with uint8 input: the new test + the hblock bench line (e2e_u8)
TAG=${1:-r02aj}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "hblock or uint8" > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest.log
timeout 900 python bench.py --config hblock --steps 30 > $OUT/bench_hb.log 2>&1; echo "bench hb $?"; tail -1 $OUT/bench_hb.log | cut -c1-200
python - <<PY
import json
d=json.loads(open("$OUT/bench_hb.log").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["e2e"]["value"], d["e2e_u8"])
PY
