#!/bin/bash
# Round-2 evidence: ncu launch lists (time + DRAM bytes) of one forward of each config, ncu --set full of the tcgen05 stem
# and of the conv kernel on the layer1 / layer4 shapes.
TAG=${1:-r02s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for CFG in resnet18 resnet50 hblock; do
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/launches_one_forward_$CFG.csv python scripts/one_forward.py --config=$CFG > $OUT/ncu_one_$CFG.log 2>&1; echo "ncu list $CFG $?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stem_tc_kernel -s 4 -c 1 \
    -o $OUT/prof_stem_tc -f python scripts/time_stem.py 256 --tc-only > $OUT/ncu_stem.log 2>&1; echo "ncu stem $?"
for L in l1 l4; do
LAYER=$L timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k 'regex:bconv_kernelILi[0-9]ELi[0-9]ELi[0-9]ELi[0-9]ELi[0-9]ELi[1234]E' -s 70 -c 1 \
    -o $OUT/prof_conv_${L}_out_cl -f python scripts/profile_layer.py --layers $L --reps 50 --fused out_cl > $OUT/ncu_conv_$L.log 2>&1; echo "ncu conv $L $?"
done
timeout 120 python -c "
import sys; sys.path.insert(0,'.')
import bnn_b200, json
from bnn_b200 import functional as BF
print(json.dumps({'popc_glanes_s': BF.ubench(0,200), 'lop3_glanes_s': BF.ubench(1,200), 'lop3_popc_iadd_gwords_s': BF.ubench(2,200), 'csa32_gwords_s': BF.ubench(3,200)}))
" > $OUT/ubench.json 2>&1; cat $OUT/ubench.json
ls -la $OUT
