#!/bin/bash
OUT=gpurun_out/${1:-tl}; mkdir -p $OUT
timeout 300 python scripts/stem_timeline.py > $OUT/timeline.jsonl 2> $OUT/timeline.err; echo "tl $?"; cat $OUT/timeline.jsonl; tail -3 $OUT/timeline.err
