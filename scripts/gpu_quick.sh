#!/bin/bash
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
grep -E "fused vs|stem kernel|resnet50|hblock" $OUT/pytest_gpu.log | head
timeout 600 python bench.py --steps 30 --warmup 3 --layers-out $OUT/layers.json > $OUT/bench.log 2>&1; tail -1 $OUT/bench.log | cut -c1-220
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/launches_one_forward.csv python scripts/one_forward.py > $OUT/ncu_one.log 2>&1; echo "ncu $?"
grep -E "stem_kernel|pack_act" $OUT/launches_one_forward.csv | cut -d, -f5,14- | head -5
