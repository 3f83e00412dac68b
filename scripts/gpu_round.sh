#!/bin/bash
# One GPU-box session: parity tests, micro-benchmarks, bench line, ncu launch list.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt
echo "== kernel tests, plain-load staging first (isolates TMA problems)" | tee $OUT/pytest_ldg.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "ldg or pack" >> $OUT/pytest_ldg.log 2>&1
echo "exit $?" >> $OUT/pytest_ldg.log
echo "== full gpu suite" | tee $OUT/pytest_gpu.log
timeout 1500 python -m pytest tests -m gpu -q >> $OUT/pytest_gpu.log 2>&1
echo "exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
echo "== micro-benchmarks"
timeout 300 python - > $OUT/ubench.json 2>$OUT/ubench.err <<'PY'
import json, sys
sys.path.insert(0, '.')
import bnn_b200
from bnn_b200 import functional as BF
names = {0: "popc_gops", 1: "lop3_gops", 2: "word_naive_gwords", 3: "word_csa32_gwords", 4: "word_csa73_gwords"}
print(json.dumps({names[k]: [BF.ubench(k, 400) for _ in range(3)] for k in names}))
PY
cat $OUT/ubench.json
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 --layers-out $OUT/layers.json > $OUT/bench.log 2>&1
tail -1 $OUT/bench.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > $OUT/bench_eager.log 2>&1
tail -1 $OUT/bench_eager.log | cut -c1-300
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 150 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
echo "ncu exit $?"
