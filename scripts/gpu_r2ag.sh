#!/bin/bash
# asynchronous shortcut staging (cp.async into per-warp landing areas) in the lean epilogue: parity suite + A/B
TAG=${1:-r02ag}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
for F in 0 64; do
  echo "== flags $F"
  python scripts/profile_layer.py --layers l1,l2,l3,l4 --reps 30 --fused out_cl --flags $F 2>&1 | tee $OUT/time_r18_out_$F.jsonl
  python scripts/profile_layer.py --layers r50_l1c3,r50_l2c3,r50_l3c3,r50_l4c3 --batch 128 --reps 30 --fused out_cl --flags $F 2>&1 | tee $OUT/time_c3_$F.jsonl
done
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/bench_r18.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18.log | cut -c1-200
timeout 600 python bench.py --config resnet50 --steps 30 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/bench_r50.log 2>&1; echo "bench r50 $?"; tail -1 $OUT/bench_r50.log | cut -c1-200
timeout 600 python bench.py --config hblock --steps 30 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/bench_hb.log 2>&1; echo "bench hb $?"; tail -1 $OUT/bench_hb.log | cut -c1-200
