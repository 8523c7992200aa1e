#!/bin/bash
# Round-2 evidence run on one B200: GPU test suite, the bench line, the launch list of one step, ncu --set full of the
# dominant kernel on three of the step's shapes.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/s1_gpu.txt 2>&1
( time timeout 780 python -m pytest tests -m gpu -x -q --durations=25 ) > gpurun_out/s1_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s1_pytest_gpu.log
( time timeout 600 python bench.py ) > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
echo "bench rc=$?" >> gpurun_out/s1_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s1_launches.csv \
    python tools/profile_step.py 32768 131072 1 > gpurun_out/s1_launches.log 2>&1
echo "launchlist rc=$?" >> gpurun_out/s1_launches.log
python tools/summarize_launches.py gpurun_out/s1_launches.csv > gpurun_out/s1_launches.md 2>&1
gzip -f gpurun_out/s1_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_i8_kernel -c 3 -f -o gpurun_out/s1_i8_full \
    python tools/i8_ncu_shapes.py > gpurun_out/s1_i8_full.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/s1_i8_full.log
ncu -i gpurun_out/s1_i8_full.ncu-rep --page raw --csv > gpurun_out/s1_i8_full_raw.csv 2>> gpurun_out/s1_i8_full.log
ls -la gpurun_out > gpurun_out/s1_ls.txt
tail -3 gpurun_out/s1_pytest_gpu.log; tail -c 600 gpurun_out/s1_bench.json; head -8 gpurun_out/s1_launches.md
