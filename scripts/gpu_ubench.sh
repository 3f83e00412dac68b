#!/bin/bash
# pipe-rate micro-benchmarks (bnn_ubench): integer modes 0-4, fp / legacy tensor path modes 5-9
TAG=${1:-ubench}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python - > $OUT/ubench.json 2>$OUT/ubench.err <<'PY'
import json, sys
sys.path.insert(0, '.')
import bnn_b200
from bnn_b200 import functional as BF
names = {0: "popc_gops", 2: "word_naive_gwords", 3: "word_csa32_gwords", 5: "mma_tf32_m16n8k8_gfma",
         6: "mma_f16_m16n8k16_gfma", 7: "mma_bf16_m16n8k16_gfma", 8: "ffma2_gfma", 9: "ffma_gfma"}
print(json.dumps({names[k]: [BF.ubench(k, 200) for _ in range(3)] for k in names}))
PY
cat $OUT/ubench.json; tail -3 $OUT/ubench.err
