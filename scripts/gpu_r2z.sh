#!/bin/bash
# programmatic dependent launch on every hot-path kernel: parity suite, then the three bench configs with and without it
TAG=${1:-r02z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
for V in pdl nopdl; do
  if [ $V = nopdl ]; then export BNN_B200_NO_PDL=1; fi
  echo "== $V"
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_r18_$V.json > $OUT/bench_r18_$V.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18_$V.log | cut -c1-200
  timeout 600 python bench.py --config resnet50 --steps 30 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_r50_$V.json > $OUT/bench_r50_$V.log 2>&1; echo "bench r50 $?"; tail -1 $OUT/bench_r50_$V.log | cut -c1-200
  timeout 600 python bench.py --config hblock --steps 30 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_hb_$V.json > $OUT/bench_hb_$V.log 2>&1; echo "bench hb $?"; tail -1 $OUT/bench_hb_$V.log | cut -c1-200
done
