"""One warm + one profiled step of the bench workload at a chosen size (for ncu; never a bench number).
    ncu ... python tools/profile_step.py N M"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from inference_tools_b200 import _lib


def synth(seed, n, d, sigma_n=0.05):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(3 * x).sum(axis=1) + rng.normal(0, sigma_n, n)
    return x, y, np.full(n, sigma_n)

n, m = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
d = 5
theta = np.array([0.2, 0.1, 1.0] + [np.log(0.3)] * d + [np.log(0.05)])
x, y, e = synth(2024, n, d)
q = np.random.default_rng(1).uniform(0, 1, (m, d))
eng = _lib.Engine(0); eng.set_data(x, y, e**2); eng.set_model([1, 2], 0)
for it in range(steps):
    l0 = eng.launch_count()
    lml, grad, info = eng.lml_grad(theta)
    eng.factor(theta)
    mu, sig = eng.predict(q)
    print("step", it, "lml", lml, "launches", eng.launch_count() - l0, flush=True)
