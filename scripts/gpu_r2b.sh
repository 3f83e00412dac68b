#!/bin/bash
# Round 2, call B: ternary / split-K tests, stem kernel timings, ncu --set full of the tcgen05 stem, launch list of one fused forward.
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "ternary or split_k or large_linear or zero" > $OUT/pytest_new.log 2>&1; echo "pytest new $?"; tail -15 $OUT/pytest_new.log
timeout 300 python scripts/time_stem.py 256 > $OUT/time_stem.json 2> $OUT/time_stem.err; echo "time_stem $?"; cat $OUT/time_stem.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stem_tc_kernel -s 4 -c 1 \
    -o $OUT/prof_stem_tc -f python scripts/time_stem.py 256 --tc-only > $OUT/ncu_stem.log 2>&1; echo "ncu stem $?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/launches_one_forward.csv python scripts/one_forward.py > $OUT/ncu_one.log 2>&1; echo "ncu list $?"
ls -la $OUT
