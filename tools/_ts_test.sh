timeout 100 python tools/i8_check2.py big 2>&1 | tail -14
for d in 0 2; do
export GPB200_GEMM_I8_DEBUG=$d
echo debug $d
timeout 150 python tools/i8_check.py i8 time 2>&1 | grep -A2 "time_"
done
