import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
KIND = {"SE": 0, "RQ": 1, "WHITE": 2, "HETERO": 3}
MEAN = {"const": 0, "linear": 1, "quadratic": 2}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names(prefix_exclude=("fit_", "linv_", "acq_")):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.startswith(prefix_exclude)]


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    d = {k: g[k] for k in g.files}
    d["comps"] = tuple(parse_term(str(c)) for c in np.atleast_1d(d["comps"]))
    d["mean"] = str(d["mean"])
    if d["x"].ndim == 1:
        d["x"] = d["x"].reshape(-1, 1)
    d["noise_var"] = d["y_err"] ** 2 if d["y_err"].size else None
    if "q" in d:
        d["q"] = d["q"].reshape(-1, d["x"].shape[1])
    return d


def parse_term(text):
    """'SE' -> 'SE';  'CP:0:SE|RQ+WHITE' -> ("CP", 0, (("SE",), ("RQ", "WHITE")))  (tests/golden/make_golden.py)"""
    if not text.startswith("CP:"):
        return text
    _, axis, regions = text.split(":")
    return ("CP", int(axis), tuple(tuple(r.split("+")) for r in regions.split("|")))


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def make_kernel(gp, comps):
    cls = {"SE": gp.SquaredExponential, "RQ": gp.RationalQuadratic, "WHITE": gp.WhiteNoise, "HETERO": gp.HeteroscedasticNoise}
    k = None
    for c in comps:
        inst = cls[c]() if isinstance(c, str) else gp.ChangePoint(kernels=[make_kernel(gp, r) for r in c[2]], axis=c[1])
        k = inst if k is None else k + inst
    return k


def make_mean(gp, mean):
    return {"const": gp.ConstantMean, "linear": gp.LinearMean, "quadratic": gp.QuadraticMean}[mean]()


def synth(seed, n, d, sigma_n=0.05):
    """Seeded synthetic regression problem of SURVEY.md section 8d."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (n, d))
    y = np.sin(3 * x).sum(axis=1) + rng.normal(0, sigma_n, n)
    return x, y, np.full(n, sigma_n)
