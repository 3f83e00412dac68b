#!/bin/bash
# what the driver does at round end, in one call: GPU suite, smoke(), default bench line
TAG=${1:-check}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke $?"; tail -1 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.log 2>&1; echo "bench $?"; tail -1 $OUT/bench.log | cut -c1-240
