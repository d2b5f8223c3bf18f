#!/bin/bash
# Multi-GPU bench exactly as the driver launches it.  Usage: gpurun --gpus N -- bash scripts/gpu_multi.sh <tag> N
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "rc=$?"; tail -c 1500 $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/ref_n$N.json 2> $OUT/ref_n$N.err
echo "ref rc=$?"; tail -c 600 $OUT/ref_n$N.json
