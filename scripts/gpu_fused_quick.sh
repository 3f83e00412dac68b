#!/bin/bash
TAG=${1:-fq}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_model.py -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest.log
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --layers-out $OUT/layers.json > $OUT/bench.log 2>&1; tail -1 $OUT/bench.log | cut -c1-200
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/launches_one_forward.csv python scripts/one_forward.py > $OUT/ncu_one.log 2>&1; echo "ncu $?"
grep -E "shortcut|pack_act" $OUT/launches_one_forward.csv | grep gpu__time | cut -d, -f5,14- | head
