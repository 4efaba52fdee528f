"""Loader for the UNMODIFIED reference modules (build container only).

/root/reference does not exist on the GPU box; nothing under tests/ (except
make_golden.py, run by hand here) may import this.
"""
import os
import sys

REF = "/root/reference"


def load():
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree not present (expected only in the build container)")
    here = os.path.dirname(os.path.abspath(__file__))
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    for p in (os.path.join(here, "_refstubs"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    # make sure we do not pick up the drop-in modules of this repo
    for n in ("simplex", "UNet", "GaussianDiffusion", "helpers", "evaluation"):
        m = sys.modules.get(n)
        if m is not None and not getattr(m, "__file__", "").startswith(REF):
            del sys.modules[n]
    import simplex, UNet, GaussianDiffusion  # noqa: E401
    assert simplex.__file__.startswith(REF), simplex.__file__
    return simplex, UNet, GaussianDiffusion
