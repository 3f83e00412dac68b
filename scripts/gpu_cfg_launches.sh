#!/bin/bash
TAG=${1:-cfgl}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for W in hblock_net resnet50_xnorpp; do
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/launches_$W.csv python scripts/one_forward_cfg.py $W > $OUT/ncu_$W.log 2>&1; echo "ncu $W $?"
done
