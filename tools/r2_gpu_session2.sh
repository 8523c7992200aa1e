#!/bin/bash
# distributed-gradient validation: single-rank tests on GPU 0, then the two-rank NCCL test when two GPUs are visible
set -u
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_dist.py -x -q --durations=8 ) > gpurun_out/s11_dist.log 2>&1
echo "rc=$?" >> gpurun_out/s11_dist.log
tail -25 gpurun_out/s11_dist.log
