#!/bin/bash
# A/B: shortcut prefetch as a dropped read-only load instead of prefetch.global.L1
TAG=${1:-r02af}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in default pfload; do
  if [ $V = pfload ]; then export BNN_B200_LIB=$PWD/binary-networks-pytorch_b200/csrc/variants/libbnn_b200_pfload.so; fi
  echo "== $V"
  python scripts/profile_layer.py --layers l1,l2,l3 --reps 30 --fused out_cl 2>&1 | tee $OUT/time_r18_out_$V.jsonl
  python scripts/profile_layer.py --layers r50_l1c3,r50_l2c3,r50_l3c3,r50_l4c3 --batch 128 --reps 30 --fused out_cl 2>&1 | tee $OUT/time_c3_$V.jsonl
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/bench_r18_$V.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18_$V.log | cut -c1-200
  timeout 600 python bench.py --config resnet50 --steps 30 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/bench_r50_$V.log 2>&1; echo "bench r50 $?"; tail -1 $OUT/bench_r50_$V.log | cut -c1-200
done
