#!/bin/bash
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "stem or amax or avgpool2 or uint8 or hblock" -s > $OUT/pytest_stem.log 2>&1; echo "pytest stem exit $?"; grep -E "^stem tc|passed|failed|Error|error" $OUT/pytest_stem.log | tail -30
timeout 300 python scripts/time_stem.py 256 > $OUT/time_stem.json 2> $OUT/time_stem.err; echo "time_stem $?"; cat $OUT/time_stem.json; tail -3 $OUT/time_stem.err
timeout 600 python scripts/flip_rate.py 64 > $OUT/flip_rate.json 2> $OUT/flip_rate.err; echo "flip $?"; cat $OUT/flip_rate.json; tail -3 $OUT/flip_rate.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stem_tc_kernel -s 4 -c 1 \
    -o $OUT/prof_stem_tc -f python scripts/time_stem.py 256 --tc-only > $OUT/ncu_stem.log 2>&1; echo "ncu stem $?"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 $OUT/pytest_gpu.log
