#!/bin/bash
OUT=gpurun_out/${1:-r02k}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -k "stem or amax or avgpool2 or uint8 or hblock" > $OUT/pytest_stem.log 2>&1; echo "pytest stem exit $?"; tail -3 $OUT/pytest_stem.log
timeout 300 python scripts/stem_roles.py > $OUT/stem_roles.json 2> $OUT/stem_roles.err; echo "roles $?"; cat $OUT/stem_roles.json; tail -3 $OUT/stem_roles.err
