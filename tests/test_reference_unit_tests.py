"""The reference's own unit tests, UNMODIFIED, against `import hepmc_b200 as hepmc`
(SURVEY.md section 4 takeaway ii; copies in tests/golden/ref_tests/ are fixtures).

The test modules do `from .. import *` / `from ..mc3 import *` relative to the package `hepmc.tests`:
the alias package below puts hepmc_b200 under the name `hepmc` and the fixture directory under
`hepmc.tests`, so the files import exactly as they do in the reference tree.  Out-of-scope updates
(adaptive Metropolis, dual averaging, NUTS, spherical HMC) are skipped by name.
"""
import importlib
import importlib.util
import os
import sys
import types
import unittest

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TESTS = os.path.join(HERE, "golden", "ref_tests")
MODULES = ["test_densities", "test_integration", "test_markov", "test_mc3"]

# SURVEY.md section 2 "OUT OF SCOPE": not part of the hot path, not built
OUT_OF_SCOPE = {
    "AdaptiveMetTest": "AdaptiveMetropolisUpdate (adaptive Metropolis) is out of scope",
    "HamiltonDualAvTest": "DualAveragingHMC is out of scope",
    "HamiltonNUTSTest": "NUTSUpdate is out of scope",
    "StaticSphericalHMCTest": "StaticSphericalHMC is out of scope",
    "DualSphericalHMCTest": "DualAveragingSphericalHMC is out of scope",
}


def _alias_package():
    """`hepmc` -> hepmc_b200 (same module objects), `hepmc.tests` -> the fixture directory."""
    import hepmc_b200
    if sys.modules.get("hepmc") is not hepmc_b200:
        sys.modules["hepmc"] = hepmc_b200
        for name, mod in list(sys.modules.items()):
            if name.startswith("hepmc_b200."):
                sys.modules["hepmc." + name[len("hepmc_b200."):]] = mod
    if "hepmc.tests" not in sys.modules:
        pkg = types.ModuleType("hepmc.tests")
        pkg.__path__ = [REF_TESTS]
        pkg.__package__ = "hepmc.tests"
        sys.modules["hepmc.tests"] = pkg
    return hepmc_b200


def _load(modname):
    _alias_package()
    full = "hepmc.tests." + modname
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(REF_TESTS, modname + ".py"))
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "hepmc.tests"
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def _collect():
    """(module, class, method) of every test the reference defines, from the files' text (no import)."""
    import ast
    cases = []
    for modname in MODULES:
        tree = ast.parse(open(os.path.join(REF_TESTS, modname + ".py")).read())
        classes = {n.name: n for n in tree.body if isinstance(n, ast.ClassDef)}

        def methods(cls, seen=()):
            out = [f.name for f in cls.body if isinstance(f, ast.FunctionDef) and f.name.startswith("test")]
            for b in cls.bases:
                bname = getattr(b, "id", None)
                if bname in classes and bname not in seen:
                    out += methods(classes[bname], seen + (bname,))
            return out
        for cname, cls in classes.items():
            if any(getattr(b, "id", None) == "TestCase" for b in cls.bases):
                for m in sorted(set(methods(cls))):
                    cases.append((modname, cname, m))
    return cases


CASES = _collect()


def test_the_fixture_files_are_the_references_files():
    """CPU: the fixtures define the reference's test inventory (and equal the checkout where it exists)."""
    assert len(CASES) == 32, CASES
    ref_dir = "/root/reference/src/hepmc/tests"
    if os.path.isdir(ref_dir):
        for modname in MODULES:
            assert open(os.path.join(ref_dir, modname + ".py")).read() == open(os.path.join(REF_TESTS, modname + ".py")).read()


@pytest.mark.gpu
@pytest.mark.parametrize("modname,cname,method", CASES, ids=["%s::%s::%s" % c for c in CASES])
def test_reference_unit_test(modname, cname, method):
    if cname in OUT_OF_SCOPE:
        pytest.skip(OUT_OF_SCOPE[cname] + " (SURVEY.md section 2)")
    H = _alias_package()
    H.set_outputs("numpy")
    H.seed(1234)
    np.random.seed(1234)
    mod = _load(modname)
    case = getattr(mod, cname)(method)
    result = unittest.TestResult()
    case.run(result)
    problems = result.errors + result.failures
    assert not problems, problems[0][1]
    assert result.testsRun == 1
