#!/bin/bash
# 3 resident CTAs for small 1x1 tiles + IMAD accumulate A/B
TAG=${1:-r02v}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
for V in default imad; do
  if [ $V = imad ]; then export BNN_B200_LIB=$PWD/binary-networks-pytorch_b200/csrc/variants/libbnn_b200_imad.so; fi
  echo "== $V"
  python scripts/profile_layer.py --layers l1,l2,l3,l4 --reps 20 --fused out_cl 2>&1 | tee $OUT/time_r18_out_$V.jsonl
  python scripts/profile_layer.py --layers l1,l2,l4,l2s --reps 20 --fused mid 2>&1 | tee $OUT/time_r18_mid_$V.jsonl
  python scripts/profile_layer.py --layers r50_l1c3,r50_l2c3,r50_l3c3,r50_l4c3 --batch 128 --reps 20 --fused out_cl 2>&1 | tee $OUT/time_c3_$V.jsonl
  python scripts/profile_layer.py --layers r50_l2c1,r50_l3c1 --batch 128 --reps 20 --fused mid 2>&1 | tee $OUT/time_c1_$V.jsonl
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_r18_$V.json > $OUT/bench_r18_$V.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18_$V.log | cut -c1-200
  timeout 600 python bench.py --config resnet50 --steps 20 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_r50_$V.json > $OUT/bench_r50_$V.log 2>&1; echo "bench r50 $?"; tail -1 $OUT/bench_r50_$V.log | cut -c1-200
done
unset BNN_B200_LIB
timeout 600 python bench.py --config resnet50 --steps 20 --warmup 5 --no-cpu-baseline --shortcut-max-cin 128 --layers-out $OUT/layers_r50_sc128.json > $OUT/bench_r50_sc128.log 2>&1; echo "bench r50 sc128 $?"; tail -1 $OUT/bench_r50_sc128.log | cut -c1-200
