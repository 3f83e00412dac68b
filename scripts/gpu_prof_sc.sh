#!/bin/bash
TAG=${1:-profsc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shortcut_kernel -s ${SKIP:-4} -c 1 \
    -o $OUT/prof_sc -f python scripts/time_shortcut.py > $OUT/ncu.log 2>&1; echo "ncu $?"
ls -la $OUT
