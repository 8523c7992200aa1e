#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/dist_grad_parity.py > gpurun_out/s14_dgp.log 2>&1; echo rc=$? >> gpurun_out/s14_dgp.log
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/s14_pytest.log 2>&1; echo rc=$? >> gpurun_out/s14_pytest.log
tail -4 gpurun_out/s14_pytest.log
GPB_BENCH_DIST_N=0 timeout 400 python bench.py > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err; echo rc=$? >> gpurun_out/s14_bench.err
