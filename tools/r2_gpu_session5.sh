#!/bin/bash
# N GPUs (2 or 4): two-rank NCCL test (N >= 2), distributed factor + gradient timing at N = 32768 (check against one GPU) and N = 131072
set -u
G=${1:-2}
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -k two_ranks ) > gpurun_out/s19_dist2.log 2>&1
echo "rc=$?" >> gpurun_out/s19_dist2.log
tail -4 gpurun_out/s19_dist2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29611 \
    tools/dist_cholesky.py --size 32768 --block 1024 --reps 1 --check --grad --out gpurun_out/s19_distgrad_N32768_g$G.json > gpurun_out/s19_distgrad_32768.log 2>&1
echo "rc=$?" >> gpurun_out/s19_distgrad_32768.log; tail -2 gpurun_out/s19_distgrad_32768.log | cut -c1-1500
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29612 \
    tools/dist_cholesky.py --size 131072 --block 2048 --reps 1 --grad --out gpurun_out/s19_distgrad_N131072_g$G.json > gpurun_out/s19_distgrad_131072.log 2>&1
echo "rc=$?" >> gpurun_out/s19_distgrad_131072.log; tail -2 gpurun_out/s19_distgrad_131072.log | cut -c1-1500
