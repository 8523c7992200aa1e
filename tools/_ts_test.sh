timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_i8.json 2> gpurun_out/bench_r1_i8.err; tail -c 600 gpurun_out/bench_r1_i8.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_i8.csv python tools/profile_step.py 32768 131072 1 > gpurun_out/profile_step.log 2>&1; tail -2 gpurun_out/profile_step.log
cat > /tmp/one.py <<'PY'
import ctypes, os, numpy as np
lib = ctypes.CDLL("inference_tools_b200/libgpb200.so")
dp = ctypes.POINTER(ctypes.c_double)
P = lambda a: a.ctypes.data_as(dp)
M=N=K=8192
rng=np.random.default_rng(0)
A=rng.standard_normal((M,K)); B=rng.standard_normal((N,K)); D=np.zeros((M,N)); ms=ctypes.c_double(0)
lib.gpb_test_gemm_impl(1, M, N, K, P(A), P(B), None, ctypes.c_double(1.0), ctypes.c_double(0.0), 0, P(D), 1, ctypes.byref(ms))
print(ms.value)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_i8_kernel -c 1 -o gpurun_out/i8_final -f python /tmp/one.py 2>&1 | tail -3
