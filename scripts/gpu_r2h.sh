#!/bin/bash
TAG=${1:-r02h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 120 ./scripts/probe/umma_rate > $OUT/umma_rate.txt 2>&1; echo "rate $?"; cat $OUT/umma_rate.txt
timeout 900 python scripts/flip_rate.py 48 hblock > $OUT/flip_hblock.json 2> $OUT/flip_hblock.err; echo "flip hb $?"; cat $OUT/flip_hblock.json; tail -3 $OUT/flip_hblock.err
timeout 900 python scripts/flip_rate.py 48 resnet50 > $OUT/flip_r50.json 2> $OUT/flip_r50.err; echo "flip r50 $?"; cat $OUT/flip_r50.json; tail -3 $OUT/flip_r50.err
