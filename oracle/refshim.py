"""Import shim for the *running reference* (test infrastructure only).

The reference (`/root/reference/src/hepmc`, pure Python/NumPy/SciPy) does not
import on NumPy 2 / without matplotlib.  This module patches the three things
it trips over and imports it from where it lies.  It is used
ONLY by `oracle/make_golden.py` (in the authoring container) and by the tests
that pin the oracle against the reference when `/root/reference` is present.
Nothing on the GPU box may depend on it: `/root/reference` does not exist there.
"""
import os
import sys
import types
import importlib

import numpy as np

REFERENCE_SRC = os.environ.get("HEPMC_REFERENCE_SRC", "/root/reference/src")


class _Inert(types.ModuleType):
    """A module whose every attribute is another inert object/callable."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _InertObj()


class _InertObj(object):
    def __call__(self, *a, **k):
        return _InertObj()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _InertObj()

    def __iter__(self):
        return iter(())


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "hepmc"))


_hepmc = None


def load():
    """Return the imported reference package `hepmc` (patched environment)."""
    global _hepmc
    if _hepmc is not None:
        return _hepmc
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_SRC)
    # (i) matplotlib is imported unconditionally (core/sampling.py:12 ->
    #     sample_plotting.py:2, stratified_volume.py:2)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors",
                     "matplotlib.cm", "matplotlib.patches", "matplotlib.gridspec",
                     "matplotlib.ticker", "mpl_toolkits",
                     "mpl_toolkits.axes_grid1"):
            sys.modules.setdefault(name, _Inert(name))
    # (ii) aliases removed from NumPy (stratified_volume.py:78,109, rambo.py:104)
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "product"):
        np.product = np.prod
    # (iii) NumPy-2 `copy=False` now raises where a copy is needed
    #       (metropolis.py:10, helper.py:25,35, gaussian.py:33,43,63)
    if not getattr(np.array, "_hepmc_shim", False):
        _orig = np.array

        def _array(*a, **k):
            if k.get("copy", True) is False:
                k["copy"] = None
            return _orig(*a, **k)

        _array._hepmc_shim = True
        np.array = _array
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    _hepmc = importlib.import_module("hepmc")
    return _hepmc
