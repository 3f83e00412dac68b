#!/bin/bash
# half-size CTAs (4 warps, twice as many resident) as tile-plan candidates: parity suite + benches + tuner log
TAG=${1:-r02al}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
BNN_B200_TUNE_LOG=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-dropin --layers-out $OUT/layers_r18.json > $OUT/bench_r18.log 2> $OUT/tune_r18.log; echo "bench r18 $?"; tail -1 $OUT/bench_r18.log | cut -c1-200
BNN_B200_TUNE_LOG=1 timeout 600 python bench.py --config resnet50 --steps 30 --warmup 5 --no-cpu-baseline --no-dropin --layers-out $OUT/layers_r50.json > $OUT/bench_r50.log 2> $OUT/tune_r50.log; echo "bench r50 $?"; tail -1 $OUT/bench_r50.log | cut -c1-200
timeout 600 python bench.py --config hblock --steps 30 --warmup 5 --no-cpu-baseline --no-dropin > $OUT/bench_hb.log 2>&1; echo "bench hb $?"; tail -1 $OUT/bench_hb.log | cut -c1-200
