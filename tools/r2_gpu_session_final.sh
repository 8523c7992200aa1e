#!/bin/bash
# Round-2 final evidence run on one B200: smoke(), GPU test suite, the bench line, the launch list of one step.
set -u
mkdir -p gpurun_out
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_smoke.log
( time timeout 780 python -m pytest tests -m gpu -x -q --durations=10 ) > gpurun_out/f_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest_gpu.log
( time timeout 600 python bench.py ) > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench rc=$?" >> gpurun_out/f_bench.err
( time timeout 400 python bench.py --impl reference --steps 1 ) > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
echo "bench ref rc=$?" >> gpurun_out/f_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches.csv \
    python tools/profile_step.py 32768 131072 1 > gpurun_out/f_launches.log 2>&1
echo "launchlist rc=$?" >> gpurun_out/f_launches.log
python tools/summarize_launches.py gpurun_out/f_launches.csv > gpurun_out/f_launches.md 2>&1
gzip -f gpurun_out/f_launches.csv
tail -3 gpurun_out/f_smoke.log; tail -3 gpurun_out/f_pytest_gpu.log; tail -c 400 gpurun_out/f_bench.json; tail -c 600 gpurun_out/f_bench_ref.json; head -6 gpurun_out/f_launches.md
