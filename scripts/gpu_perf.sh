#!/bin/bash
# Perf-focused GPU session: quick parity re-check, isolated layers, bench, ncu full capture of conv kernels.
TAG=${1:-perf}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 300 python scripts/profile_layer.py --layers l1,l2s,l2,l2d,l3s,l3,l3d,l4s,l4,l4d --reps 10 | tee $OUT/layers_iso.jsonl
timeout 300 python scripts/profile_layer.py --layers l1,l2,l3,l4 --reps 10 --flags 2 | tee $OUT/layers_iso_nocsa.jsonl
timeout 900 python bench.py --steps 30 --warmup 3 --layers-out $OUT/layers.json > $OUT/bench.log 2>&1
tail -1 $OUT/bench.log | cut -c1-400
if [ "${NCU:-1}" = "1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bconv_kernel -s 1 -c 1 \
    -o $OUT/prof_l1 -f python scripts/profile_layer.py --layers l1 --reps 2 > $OUT/ncu_l1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bconv_kernel -s 1 -c 1 \
    -o $OUT/prof_l4 -f python scripts/profile_layer.py --layers l4 --reps 2 > $OUT/ncu_l4.log 2>&1
ls -la $OUT/*.ncu-rep
fi
