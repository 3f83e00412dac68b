#!/bin/bash
# Round-2 closing evidence on one B200: GPU parity suite, smoke(), the three bench lines (with the reference CPU arm as
# cpu_baseline), ncu launch lists of one forward per config, ncu --set full of the conv kernel (layer1 / layer4 shapes,
# fused NHWC epilogue) and of the tcgen05 stem.
TAG=${1:-r02final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke $?"; tail -1 $OUT/smoke.log
timeout 900 python bench.py --layers-out $OUT/layers_r18.json > $OUT/bench_r18.log 2>&1; echo "bench r18 $?"; tail -1 $OUT/bench_r18.log | cut -c1-300
timeout 900 python bench.py --config resnet50 --steps 30 --layers-out $OUT/layers_r50.json > $OUT/bench_r50.log 2>&1; echo "bench r50 $?"; tail -1 $OUT/bench_r50.log | cut -c1-200
timeout 900 python bench.py --config hblock --steps 30 --layers-out $OUT/layers_hb.json > $OUT/bench_hb.log 2>&1; echo "bench hb $?"; tail -1 $OUT/bench_hb.log | cut -c1-200
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_arm.log 2>&1; echo "reference arm $?"; tail -1 $OUT/bench_reference_arm.log | cut -c1-300
for CFG in resnet18 resnet50 hblock; do
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $OUT/launches_one_forward_$CFG.csv python scripts/one_forward.py --config=$CFG > $OUT/ncu_one_$CFG.log 2>&1; echo "ncu list $CFG $?"
done
for L in l1 l4; do
for F in out_cl mid; do
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k 'regex:bconv_kernelILi[0-9]ELi[0-9]ELi[0-9]ELi[0-9]ELi[0-9]ELi[1234]E' -s 120 -c 1 \
    -o $OUT/prof_conv_${L}_$F -f python scripts/profile_layer.py --layers $L --reps 200 --fused $F > $OUT/ncu_conv_${L}_$F.log 2>&1; echo "ncu conv $L $F $?"
done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stem_tc_kernel -s 4 -c 1 \
    -o $OUT/prof_stem_tc -f python scripts/time_stem.py 256 --tc-only > $OUT/ncu_stem.log 2>&1; echo "ncu stem $?"
ls $OUT | wc -l
