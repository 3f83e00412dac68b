#!/bin/bash
TAG=${1:-fused}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
grep -E "fused vs|stem kernel|passed|failed|Error" $OUT/pytest_gpu.log | tail -8
for mode in "" mid out_cl; do
timeout 300 python scripts/profile_layer.py --layers l1,l2s,l2,l3s,l3,l4s,l4 --reps 10 --fused "$mode" | tee -a $OUT/layers_iso.jsonl
done
timeout 900 python bench.py --steps 30 --warmup 3 --layers-out $OUT/layers.json > $OUT/bench.log 2>&1
tail -1 $OUT/bench.log | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv \
    --log-file $OUT/launches_fused.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $OUT/ncu_bench.log 2>&1
echo "ncu exit $?"
if [ "${NCU:-0}" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_kernel -s 1 -c 1 \
    -o $OUT/prof_stem -f python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > $OUT/ncu_stem.log 2>&1
fi
