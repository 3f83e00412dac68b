#!/bin/bash
TAG=${1:-profstem}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stem_mma -s ${SKIP:-3} -c 1 \
    -o $OUT/prof_stem_mma -f python scripts/time_stem.py ${BS:-64} > $OUT/ncu_stem.log 2>&1; echo "ncu $?"
ls -la $OUT
