#!/bin/bash
OUT=gpurun_out/${1:-dbg}; mkdir -p $OUT
timeout 600 python scripts/debug_engine2.py > $OUT/debug_engine2.log 2>&1; echo "dbg $?"; cat $OUT/debug_engine2.log | cut -c1-600
