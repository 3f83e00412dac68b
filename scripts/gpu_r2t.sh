#!/bin/bash
# ResNet-50 1x1 layers in isolation: event timing + ncu --set full of the conv kernel per shape
TAG=${1:-r02t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python scripts/profile_layer.py --layers r50_l1c3,r50_l2c3,r50_l3c3,r50_l4c3 --batch 128 --reps 20 --fused out_cl > $OUT/time_c3.jsonl 2>&1
python scripts/profile_layer.py --layers r50_l2c1,r50_l3c1 --batch 128 --reps 20 --fused mid > $OUT/time_c1.jsonl 2>&1
cat $OUT/time_c3.jsonl $OUT/time_c1.jsonl
for L in r50_l1c3 r50_l2c3 r50_l3c3; do
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k 'regex:bconv_kernelILi[0-9]ELi[0-9]ELi[0-9]ELi[0-9]ELi[0-9]ELi[1234]E' -s 70 -c 1 \
    -o $OUT/prof_${L} -f python scripts/profile_layer.py --layers $L --batch 128 --reps 50 --fused out_cl > $OUT/ncu_$L.log 2>&1; echo "ncu $L $?"
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k 'regex:bconv_kernelILi[0-9]ELi[0-9]ELi[0-9]ELi[0-9]ELi[0-9]ELi[1234]E' -s 70 -c 1 \
    -o $OUT/prof_r50_l3c1 -f python scripts/profile_layer.py --layers r50_l3c1 --batch 128 --reps 50 --fused mid > $OUT/ncu_l3c1.log 2>&1; echo "ncu l3c1 $?"
ls -la $OUT
