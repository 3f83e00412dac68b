#!/bin/bash
# multi-GPU check: bench.py under torchrun exactly as the driver launches it
N=${1:-2}
TAG=${2:-multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 3 > $OUT/bench_n$N.log 2>&1
echo "exit $?"; tail -1 $OUT/bench_n$N.log | cut -c1-900
