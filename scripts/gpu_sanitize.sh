#!/bin/bash
# compute-sanitizer over a subset of the GPU parity tests (memcheck on all kernels, racecheck + synccheck on the
# shared-memory heavy ones).  Slow: small cases only.
TAG=${1:-san}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
SEL='stem or shortcut or avgpool or (fused_epilogue and tma) or pack_activations or known_answer'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file $OUT/memcheck.log \
    python -m pytest tests/test_gpu_fused.py tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL" > $OUT/pytest_memcheck.log 2>&1
echo "memcheck exit $?"; tail -2 $OUT/pytest_memcheck.log; grep -E "ERROR SUMMARY|Invalid|misaligned" $OUT/memcheck.log | tail -3
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 99 --log-file $OUT/racecheck.log \
    python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "stem_mma_kernel or shortcut_kernel_bit or (fused_epilogue and tma and (basic or pre))" > $OUT/pytest_racecheck.log 2>&1
echo "racecheck exit $?"; tail -2 $OUT/pytest_racecheck.log; grep -E "RACECHECK SUMMARY|hazard" $OUT/racecheck.log | tail -3
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 99 --log-file $OUT/synccheck.log \
    python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "stem_mma_kernel or shortcut_kernel_bit" > $OUT/pytest_synccheck.log 2>&1
echo "synccheck exit $?"; tail -2 $OUT/pytest_synccheck.log; grep -E "ERROR SUMMARY" $OUT/synccheck.log | tail -2
