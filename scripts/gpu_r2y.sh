#!/bin/bash
# persistent plans accepted only when clearly faster: parity suite + the three bench configs
TAG=${1:-r02y}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_r18.json > $OUT/bench_r18.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18.log | cut -c1-200
timeout 600 python bench.py --config resnet50 --steps 20 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_r50.json > $OUT/bench_r50.log 2>&1; echo "bench r50 $?"; tail -1 $OUT/bench_r50.log | cut -c1-200
timeout 600 python bench.py --config hblock --steps 20 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_hb.json > $OUT/bench_hb.log 2>&1; echo "bench hb $?"; tail -1 $OUT/bench_hb.log | cut -c1-200
