#!/bin/bash
# racecheck without the mbarrier-synchronised tcgen05 stem (its producer/consumer rings are reported as hazards because
# racecheck does not model mbarrier waits): lean conv epilogue, shortcut kernel, dense pack, pooled pack
TAG=${1:-san4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --log-file $OUT/racecheck.log \
    python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "(fused_epilogue and lean and tma) or shortcut_kernel_bit or dense or avgpool2_pack" > $OUT/pytest_racecheck.log 2>&1
echo "racecheck exit $?"; tail -1 $OUT/pytest_racecheck.log; grep -E "RACECHECK SUMMARY" $OUT/racecheck.log | tail -1; grep -c "Race reported\|hazard" $OUT/racecheck.log
