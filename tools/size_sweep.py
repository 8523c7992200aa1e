"""LML-gradient phase timings vs N (gpurun)."""
import sys, time, numpy as np
sys.path.insert(0, ".")
from inference_tools_b200 import _lib
from oracle.cpu_reference import synth
for n in (256, 1024, 2048, 4096, 8192, 16384):
    x, y, e = synth(1, n, 3); th = np.array([0.3, 0.1] + [np.log(0.3)] * 3)
    eng = _lib.Engine(0); eng.set_data(x, y, e**2); eng.set_model([0], 0)
    for it in range(4):
        t0 = time.perf_counter(); lml, g, info = eng.lml_grad(th); t1 = time.perf_counter(); tg = eng.timers()
    t2 = time.perf_counter(); eng.lml(th); tl = time.perf_counter() - t2
    print(n, "lml_grad wall ms", round((t1 - t0) * 1e3, 2), {k: round(v, 3) for k, v in tg.items()}, "lml wall ms", round(tl * 1e3, 2), flush=True)
    eng.close()
