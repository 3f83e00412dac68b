#!/bin/bash
# shortcuts on a second stream: parity suite, then the benches with and without the overlap
TAG=${1:-r02ac}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
for V in overlap serial; do
  X=""; if [ $V = serial ]; then X="--no-overlap"; fi
  echo "== $V"
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-dropin $X --layers-out $OUT/layers_r18_$V.json > $OUT/bench_r18_$V.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18_$V.log | cut -c1-200
  timeout 600 python bench.py --config resnet50 --steps 30 --warmup 5 --no-cpu-baseline --no-dropin $X --layers-out $OUT/layers_r50_$V.json > $OUT/bench_r50_$V.log 2>&1; echo "bench r50 $?"; tail -1 $OUT/bench_r50_$V.log | cut -c1-200
done
