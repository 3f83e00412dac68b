#!/bin/bash
# lean-epilogue rework: parity suite, isolated layers, the three bench configs
TAG=${1:-r02u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
python scripts/profile_layer.py --layers l1,l2,l3,l4 --reps 20 --fused out_cl > $OUT/time_r18_out.jsonl 2>&1
python scripts/profile_layer.py --layers l1,l2,l3,l4,l2s,l3s,l4s --reps 20 --fused mid > $OUT/time_r18_mid.jsonl 2>&1
python scripts/profile_layer.py --layers r50_l1c3,r50_l2c3,r50_l3c3,r50_l4c3 --batch 128 --reps 20 --fused out_cl > $OUT/time_c3.jsonl 2>&1
python scripts/profile_layer.py --layers r50_l2c1,r50_l3c1 --batch 128 --reps 20 --fused mid > $OUT/time_c1.jsonl 2>&1
cat $OUT/time_*.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 --layers-out $OUT/layers_r18.json > $OUT/bench_r18.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18.log | cut -c1-200
timeout 600 python bench.py --config resnet50 --steps 20 --warmup 5 --layers-out $OUT/layers_r50.json > $OUT/bench_r50.log 2>&1; echo "bench r50 $?"; tail -1 $OUT/bench_r50.log | cut -c1-200
timeout 600 python bench.py --config hblock --steps 20 --warmup 5 --layers-out $OUT/layers_hb.json > $OUT/bench_hb.log 2>&1; echo "bench hb $?"; tail -1 $OUT/bench_hb.log | cut -c1-200
