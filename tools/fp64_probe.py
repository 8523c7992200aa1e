"""FP64 roofline denominators on the B200 box: cuBLAS DGEMM and cuSOLVER potrf through torch
(library kernels = the "kernel to beat"), plus the raw DMMA/DFMA issue-rate
probe (tools/fp64_probe.cu).  Writes gpurun_out/fp64_peak.json."""
import json, os, subprocess, sys, time
import torch

out = {}
dev = torch.device("cuda:0")
out["gpu"] = torch.cuda.get_device_name(0)
out["host_cpus"] = os.cpu_count()
try:
    with open("/proc/meminfo") as f:
        out["host_mem_gb"] = int(f.readline().split()[1]) / 1e6
except Exception:
    pass

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best

for n in (4096, 8192, 16384):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    t = timed(lambda: torch.mm(a, b.T))
    out[f"cublas_dgemm_nt_{n}_tflops"] = 2 * n**3 / t / 1e12
    t = timed(lambda: torch.mm(a, b))
    out[f"cublas_dgemm_nn_{n}_tflops"] = 2 * n**3 / t / 1e12
    del a, b
# sustained dgemm ~4 s
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); reps = 0; t0 = time.time()
while time.time() - t0 < 4.0:
    for _ in range(10):
        torch.mm(a, b.T)
    reps += 10; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
out["cublas_dgemm_nt_8192_tflops_sustained"] = reps * 2 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
del a, b

for n in (8192, 16384, 32768):
    x = torch.rand(n, 3, dtype=torch.float64, device=dev)
    d2 = torch.cdist(x, x) ** 2
    k = torch.exp(-0.5 * d2 / 0.09) + 0.0025 * torch.eye(n, dtype=torch.float64, device=dev)
    del d2
    t = timed(lambda: torch.linalg.cholesky(k), reps=2)
    out[f"cusolver_potrf_{n}_s"] = t
    out[f"cusolver_potrf_{n}_tflops"] = n**3 / 3 / t / 1e12
    del k, x
    torch.cuda.empty_cache()

exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fp64_probe")
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                        "--format=csv,noheader", "-lms", "200"], stdout=subprocess.PIPE, text=True)
r = subprocess.run([exe], capture_output=True, text=True)
smi.terminate()
clk = smi.communicate()[0].strip().splitlines()
out["probe_raw"] = r.stdout.strip()
try:
    out["probe"] = json.loads(r.stdout.strip().splitlines()[-1])
except Exception as e:
    out["probe_err"] = repr(e) + r.stderr
sm = sorted(int(l.split(",")[0].split()[0]) for l in clk if l and l[0].isdigit())
out["clocks_during_probe"] = {"sm_mhz_median": sm[len(sm) // 2] if sm else None, "sm_mhz_min": sm[0] if sm else None,
                              "sm_mhz_max": sm[-1] if sm else None, "samples": len(sm), "last": clk[-3:] if clk else []}
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/fp64_peak.json", "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
