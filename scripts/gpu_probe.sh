#!/bin/bash
OUT=gpurun_out/${1:-probe}; mkdir -p $OUT
timeout 120 ./scripts/probe/umma_row > $OUT/umma_row.txt 2>&1; echo "row $?"; cat $OUT/umma_row.txt
