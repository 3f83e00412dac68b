#!/bin/bash
TAG=${1:-r02i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "stem or amax or avgpool2 or uint8 or hblock" -s > $OUT/pytest_stem.log 2>&1; echo "pytest stem exit $?"; grep -E "uint8 engine|passed|failed|Error|error" $OUT/pytest_stem.log | tail -30
timeout 300 python scripts/time_stem.py 256 > $OUT/time_stem.json 2> $OUT/time_stem.err; echo "time_stem $?"; cat $OUT/time_stem.json; tail -3 $OUT/time_stem.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stem_tc_kernel -s 4 -c 1 \
    -o $OUT/prof_stem_tc -f python scripts/time_stem.py 256 --tc-only > $OUT/ncu_stem.log 2>&1; echo "ncu stem $?"
