#!/bin/bash
# compute-sanitizer over the kernel-level parity tests (all conv / pack / stem / shortcut cases, no whole models)
TAG=${1:-san2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SEL='not engine and not pipeline and not hblock_net and not resnet50 and not module_api and not weight_cache'
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file $OUT/memcheck.log \
    python -m pytest tests/test_gpu_fused.py tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL" > $OUT/pytest_memcheck.log 2>&1
echo "memcheck exit $?"; tail -1 $OUT/pytest_memcheck.log; grep -E "ERROR SUMMARY" $OUT/memcheck.log | tail -1
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 99 --log-file $OUT/racecheck.log \
    python -m pytest tests/test_gpu_fused.py tests/test_gpu_kernels.py -m gpu -q -x -k "(fused_epilogue and tma and lean) or shortcut_kernel_bit or dense" > $OUT/pytest_racecheck.log 2>&1
echo "racecheck exit $?"; tail -1 $OUT/pytest_racecheck.log; grep -E "RACECHECK SUMMARY" $OUT/racecheck.log | tail -1
