#!/bin/bash
# two GPUs: the two-rank NCCL test of the distributed regressor (gradient included), then gradient timing at N = 16384 / 32768
set -u
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -k two_ranks ) > gpurun_out/s12_dist2.log 2>&1
echo "rc=$?" >> gpurun_out/s12_dist2.log
tail -5 gpurun_out/s12_dist2.log
for n in 16384 32768; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    tools/dist_cholesky.py --size $n --block 1024 --reps 1 --check --grad --out gpurun_out/s12_distgrad_N${n}_g2.json > gpurun_out/s12_distgrad_${n}.log 2>&1
echo "rc=$?" >> gpurun_out/s12_distgrad_${n}.log
tail -3 gpurun_out/s12_distgrad_${n}.log | cut -c1-1200
done
