#!/bin/bash
# Round 2, call A: tcgen05 descriptor probe, tcgen05 stem tests, full GPU parity suite (incl. the forced-plan sweep),
# new bench.py flow on all three configs, reference arm with the byte-compiled reference.
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 60 ./scripts/probe/umma_probe > $OUT/probe.txt 2>&1; echo "probe exit $?"; cat $OUT/probe.txt
timeout 400 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "stem or amax" -s > $OUT/pytest_stem.log 2>&1; echo "pytest stem exit $?"; grep -E "^stem|passed|failed|Error|error" $OUT/pytest_stem.log | tail -40
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -30 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --layers-out $OUT/layers_r18.json > $OUT/bench_r18.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18.log | cut -c1-600
timeout 600 python bench.py --steps 20 --warmup 5 --stem mma --no-dropin --no-cpu-baseline > $OUT/bench_r18_mma.log 2>&1; echo "bench r18 mma $?"; tail -1 $OUT/bench_r18_mma.log | cut -c1-300
timeout 600 python bench.py --config resnet50 --steps 20 --warmup 5 --layers-out $OUT/layers_r50.json > $OUT/bench_r50.log 2>&1; echo "bench r50 $?"; tail -1 $OUT/bench_r50.log | cut -c1-400
timeout 600 python bench.py --config hblock --steps 20 --warmup 5 --layers-out $OUT/layers_hb.json > $OUT/bench_hb.log 2>&1; echo "bench hb $?"; tail -1 $OUT/bench_hb.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.log 2>&1; echo "ref $?"; tail -1 $OUT/bench_ref.log | cut -c1-400
