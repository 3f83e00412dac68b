#!/bin/bash
# N = 2: both arms as the driver launches them, plus the gloo-free NCCL parity of the gathered logits inside bench.py
TAG=${1:-r02r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_2gpu.log 2> $OUT/bench_2gpu.err; echo "bench 2gpu $?"; tail -1 $OUT/bench_2gpu.log | cut -c1-400; tail -5 $OUT/bench_2gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $OUT/bench_ref_2gpu.log 2> $OUT/bench_ref_2gpu.err; echo "ref 2gpu $?"; tail -1 $OUT/bench_ref_2gpu.log | cut -c1-300
