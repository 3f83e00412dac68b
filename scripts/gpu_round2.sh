#!/bin/bash
# Full round check: parity suite, smoke, bench (both arms), launch list + DRAM traffic of one fused forward.
TAG=${1:-round}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1; nproc >> $OUT/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --layers-out $OUT/layers.json > $OUT/bench.log 2>&1; tail -1 $OUT/bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.log 2>&1; tail -1 $OUT/bench_ref.log | cut -c1-300
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/launches_one_forward.csv python scripts/one_forward.py > $OUT/ncu_one.log 2>&1; echo "ncu $?"
