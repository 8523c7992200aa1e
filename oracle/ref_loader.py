"""Loader for the UNMODIFIED reference package -- TEST / BENCH INFRASTRUCTURE, never imported by the product.

`__graft_entry__.build()` installs the reference into baseline/_ref/ (git-ignored, travels to the GPU box with the
gpurun snapshot): `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`.
The reference's own packaging lists only the top-level package (`pyproject.toml: [tool.setuptools] packages =
["inference"]`), so the wheel pip builds carries no sub-packages; build() completes the install by copying the missing
sub-package directories (gp/, mcmc/, pdf/, approx/) verbatim from the same source tree.  Nothing under baseline/_ref is
edited.  matplotlib is absent from this image: it is stubbed at import time -- the GP path never
calls it.
"""
from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]


def reference_dir():
    for base in REF_DIRS:
        if os.path.isdir(os.path.join(base, "inference", "gp")):
            return base
    return None


def load_reference_gp():
    """The reference's `inference.gp` module (unmodified source), or None when no copy is available on this host."""
    base = reference_dir()
    if base is None:
        return None
    if "inference" not in sys.modules or not getattr(sys.modules["inference"], "__gpb_ref__", False):
        pkg = types.ModuleType("inference")            # skips inference/__init__.py (needs package metadata)
        pkg.__path__ = [os.path.join(base, "inference")]
        pkg.__gpb_ref__ = True
        sys.modules["inference"] = pkg
        for name in ("matplotlib", "matplotlib.pyplot"):
            sys.modules.setdefault(name, types.ModuleType(name))
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        if not hasattr(sys.modules["matplotlib"], "__path__"):
            sys.modules["matplotlib"].__path__ = []
    import inference.gp as ref_gp
    return ref_gp
