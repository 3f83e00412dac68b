#!/bin/bash
# N = 8: both arms as the driver launches them, plus the gloo-free NCCL parity of the gathered logits inside bench.py
TAG=${1:-r02r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_8gpu.log 2> $OUT/bench_8gpu.err; echo "bench 8gpu $?"; tail -1 $OUT/bench_8gpu.log | cut -c1-400; tail -5 $OUT/bench_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29524 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > $OUT/bench_ref_8gpu.log 2> $OUT/bench_ref_8gpu.err; echo "ref 8gpu $?"; tail -1 $OUT/bench_ref_8gpu.log | cut -c1-300
