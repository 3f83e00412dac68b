#!/bin/bash
# round 2: compute-sanitizer over the kernel-level parity tests (conv incl. the plan sweep on one geometry, pack, stems,
# shortcut) -- memcheck everywhere, racecheck + synccheck on the lean epilogue / shortcut / tcgen05 stem cases
TAG=${1:-san3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
SEL='not engine and not pipeline and not hblock_net and not resnet50 and not module_api and not weight_cache and not second_stream and not uint8_images'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file $OUT/memcheck.log \
    python -m pytest tests/test_gpu_fused.py tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL" > $OUT/pytest_memcheck.log 2>&1
echo "memcheck exit $?"; tail -1 $OUT/pytest_memcheck.log; grep -E "ERROR SUMMARY" $OUT/memcheck.log | tail -1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file $OUT/memcheck_plans.log \
    python -m pytest tests/test_gpu_plans.py -m gpu -q -x -k "k3s1 or k1s1" > $OUT/pytest_memcheck_plans.log 2>&1
echo "memcheck plans exit $?"; tail -1 $OUT/pytest_memcheck_plans.log; grep -E "ERROR SUMMARY" $OUT/memcheck_plans.log | tail -1
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 99 --log-file $OUT/racecheck.log \
    python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "(fused_epilogue and lean) or shortcut_kernel_bit or dense or stem_tc_matches" > $OUT/pytest_racecheck.log 2>&1
echo "racecheck exit $?"; tail -1 $OUT/pytest_racecheck.log; grep -E "RACECHECK SUMMARY" $OUT/racecheck.log | tail -1
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 99 --log-file $OUT/synccheck.log \
    python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "(fused_epilogue and lean) or shortcut_kernel_bit or stem_tc_matches" > $OUT/pytest_synccheck.log 2>&1
echo "synccheck exit $?"; tail -1 $OUT/pytest_synccheck.log; grep -E "ERROR SUMMARY" $OUT/synccheck.log | tail -1
