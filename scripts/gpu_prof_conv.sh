#!/bin/bash
# ncu --set full of the FUSED conv kernel (EPI 1 instances) on one isolated layer: LAYER (l1..l4), MODES (mid out_cl)
TAG=${1:-profconv}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for MODE in ${MODES:-out_cl mid}; do
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k 'regex:bconv_kernelILi[0-9]ELi[0-9]ELi[0-9]ELi[0-9]ELi[0-9]ELi[1234]E' -s 70 -c 1 \
    -o $OUT/prof_${LAYER:-l1}_$MODE -f python scripts/profile_layer.py --layers ${LAYER:-l1} --reps 50 --fused $MODE > $OUT/ncu_$MODE.log 2>&1; echo "ncu $?"
done
ls -la $OUT
