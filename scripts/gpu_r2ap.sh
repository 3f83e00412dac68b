#!/bin/bash
# A/B: C <= 2 conv instances compiled for 96 registers (5 resident 4-warp CTAs = 20 warps per SM instead of 16)
TAG=${1:-r02aq}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in default c2r80; do
  if [ $V = c2r80 ]; then export BNN_B200_LIB=$PWD/binary-networks-pytorch_b200/csrc/variants/libbnn_b200_c2r80.so; fi
  echo "== $V"
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-dropin --layers-out $OUT/layers_r18_$V.json > $OUT/bench_r18_$V.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18_$V.log | cut -c1-200
  timeout 600 python bench.py --config resnet50 --steps 30 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/bench_r50_$V.log 2>&1; echo "bench r50 $?"; tail -1 $OUT/bench_r50_$V.log | cut -c1-200
done
timeout 600 python -m pytest tests/test_gpu_plans.py -m gpu -q 2>&1 | tail -1
