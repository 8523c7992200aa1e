for d in 0 1 2; do
export GPB200_GEMM_I8_DEBUG=$d
echo debug $d
timeout 150 python tools/i8_check.py i8 time 2>&1 | grep -A1 "time_8192\|time_16384x16384x2048"
done
