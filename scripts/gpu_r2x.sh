#!/bin/bash
# persistent conv plans (two-deep window ring): parity sweep, isolated layers, benches; ncu of the shortcut kernel on R50 shapes
TAG=${1:-r02x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
python scripts/profile_layer.py --layers l1,l2,l3,l4 --reps 20 --fused out_cl 2>&1 | tee $OUT/time_r18_out.jsonl
python scripts/profile_layer.py --layers l1,l2,l4,l2s --reps 20 --fused mid 2>&1 | tee $OUT/time_r18_mid.jsonl
python scripts/profile_layer.py --layers r50_l1c3,r50_l2c3,r50_l3c3,r50_l4c3 --batch 128 --reps 20 --fused out_cl 2>&1 | tee $OUT/time_c3.jsonl
python scripts/profile_layer.py --layers r50_l2c1,r50_l3c1 --batch 128 --reps 20 --fused mid 2>&1 | tee $OUT/time_c1.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_r18.json > $OUT/bench_r18.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18.log | cut -c1-200
timeout 600 python bench.py --config resnet50 --steps 20 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_r50.json > $OUT/bench_r50.log 2>&1; echo "bench r50 $?"; tail -1 $OUT/bench_r50.log | cut -c1-200
timeout 600 python bench.py --config hblock --steps 20 --warmup 5 --no-cpu-baseline --layers-out $OUT/layers_hb.json > $OUT/bench_hb.log 2>&1; echo "bench hb $?"; tail -1 $OUT/bench_hb.log | cut -c1-200
python scripts/time_shortcut.py --r50 2>&1 | tee $OUT/time_shortcut_r50.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shortcut_kernel -s 210 -c 1 -o $OUT/prof_sc_l2 -f python scripts/time_shortcut.py --r50 > $OUT/ncu_sc_l2.log 2>&1; echo "ncu sc l2 $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shortcut_kernel -s 630 -c 1 -o $OUT/prof_sc_l4 -f python scripts/time_shortcut.py --r50 > $OUT/ncu_sc_l4.log 2>&1; echo "ncu sc l4 $?"
