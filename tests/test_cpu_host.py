"""CPU-only checks: the C-ABI library loads and exports every declared symbol, host-side
logic (shape contract, grid state, slot partition) and the loud failure without a GPU."""
import os
import re
import ctypes

import numpy as np
import pytest
import torch

import hepmc_b200 as H
from hepmc_b200 import _lib, parallel
from hepmc_b200.core.util import interpret_array, damped_update, assure_2d
from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hepmc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hepmc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "missing symbol " + name
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.load().hepmc_abi_version() == 2


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.DensityBlock) == 8 + 3 * 32 * 8 + 4 * 8
    assert ctypes.sizeof(_lib.ElmBlock) == 8 + 3 * 8 + 8


def test_host_only_entry_points():
    lib = _lib.load()
    assert lib.hepmc_rambo_pdf(4, 100.0) == pytest.approx(3.0961473055871514e-08, rel=1e-14)
    for n, chunk in ((1000, 2048), (1 << 21, 8192), (10 ** 9, 32768)):
        assert lib.hepmc_default_chunk(n) == chunk
    assert lib.hepmc_vegas_workspace_bytes(16, 15, 10 ** 9) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_fails_loudly_without_gpu():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.densities.Camel(2).pdf(np.zeros((3, 2)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.PlainMC(2)(H.densities.Camel(2), 100)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.VegasMC(2, 4)(H.densities.Camel(2), 100, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.DefaultMetropolis(2, H.densities.Banana(2), H.proposals.Gaussian(2, cov=10.)).sample(10, [0., 10.])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.phase_space.Rambo(100., 4).map(np.full((2, 16), 0.5))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hepmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_interpret_array_contract():
    """core/util/helper.py:34-53"""
    assert interpret_array(np.zeros((5, 3)), 3).shape == (5, 3)
    assert interpret_array(np.zeros(5), 1).shape == (5, 1)
    assert interpret_array(np.zeros(3), 3).shape == (1, 3)
    assert interpret_array(0.5, 1).shape == (1, 1)
    with pytest.raises(RuntimeError):
        interpret_array(np.zeros(4), 3)
    with pytest.raises(RuntimeWarning):
        interpret_array(np.zeros((4, 2)), 3)
    assert interpret_array(torch.zeros(3), 3).shape == (1, 3)
    assert assure_2d(np.arange(4)).shape == (4, 1)
    with pytest.raises(RuntimeError):
        assure_2d(np.zeros((2, 2, 2)))


def test_damped_update():
    old, new = np.array([1.0, 2.0]), np.array([3.0, 1.0])
    np.testing.assert_allclose(damped_update(old, new, 3, 0), new * 3 / 4 + old * 1 / 4)
    np.testing.assert_allclose(damped_update(old, new, 3, 5), new * 3 / 9 + old * 6 / 9)


def test_grid_volumes_host_state():
    g = H.GridVolumes(ndim=3, divisions=4)
    assert g.divisions == 4 and g.partition_count == 64.0
    np.testing.assert_array_equal(g.bounds[1], np.linspace(0, 1, 5))
    np.testing.assert_allclose(g.sizes[2], np.full(4, 0.25))
    g.sizes = [np.array([0.1, 0.2, 0.3, 0.4])] * 3
    assert g.bounds[0][0] == 0.0                      # App. B2
    np.testing.assert_allclose(g.bounds[0], [0, 0.1, 0.3, 0.6, 1.0])
    g.reset()
    np.testing.assert_array_equal(g.bounds[0], np.linspace(0, 1, 5))
    # no int64 overflow of the partition count (App. B1)
    assert H.GridVolumes(ndim=16, divisions=50).partition_count == 50.0 ** 16
    mc = H.VegasMC(2, divisions=5)
    idx = np.array([[0, 1], [4, 2]])
    np.testing.assert_allclose(mc.weights_at(idx), [1.0, 1.0])


def test_density_blocks_follow_scipy_constants():
    c = H.densities.Camel(4)
    blk = c._block()
    assert blk.kind == _lib.CAMEL and blk.ndim == 4
    assert blk.s0 == np.sqrt(1.0 / 0.005) or abs(blk.s0 - np.sqrt(1 / (0.1 ** 2 / 2))) < 1e-15
    assert blk.s1 == pytest.approx(4 * np.log(2 * np.pi) + 4 * np.log(0.1 ** 2 / 2), rel=1e-15)
    g = H.densities.Gaussian(3, mu=[1, 2, 3], cov=[0.5, 1.0, 2.0])
    b = g._block()
    assert [b.a[i] for i in range(3)] == [1, 2, 3]
    assert b.c[2] == 0.5
    # the cov setter multiplies by the identity (gaussian.py:73-76): always diagonal
    assert np.count_nonzero(H.densities.Gaussian(2, cov=np.array([[1, .1], [.1, 1]])).cov) == 2


def test_slot_partition_is_invariant():
    for n_total in (100, 2048 * 7 + 5, 10 ** 6, 10 ** 9):
        chunk = _lib.default_chunk(n_total)
        n_chunks = (n_total + chunk - 1) // chunk
        ref = parallel.slot_begins(n_chunks)
        assert ref[0] == 0 and ref[-1] == n_chunks
        for ws in (1, 2, 4, 8):
            cover, slots = 0, []
            for r in range(ws):
                first, n_local, begins = parallel.local_span(n_total, chunk, r, ws)
                assert first % chunk == 0 and first == cover or n_local == 0
                cover = first + n_local if n_local else cover
                c0 = first // chunk
                slots += [c0 + b for b in begins[:-1]] if n_local or True else []
            assert cover == n_total
    with pytest.raises(RuntimeError):
        parallel.slot_span(0, 3)


def test_rng_streams():
    H.seed(7)
    assert H.rng.get_seed() == 7 and H.rng.next_stream() == 0 and H.rng.next_stream(10) == 1
    assert H.rng.next_stream() == 11
    H.seed(1234)


def test_stat_tools_host_logic_matches_numpy():
    """Bin edges and percentiles prepared on the host for the statistics kernels are NumPy's
    (np.histogramdd / np.percentile 'linear'), checked on CPU tensors."""
    import torch
    from hepmc_b200.core.util import stat_tools
    rs = np.random.RandomState(3)
    data = rs.randn(501, 3) * [1., 0.1, 7.]
    x = torch.from_numpy(data)
    for bins, rng_ in ((7, None), ((3, 4, 5), None), (6, ((-1, 1), (-0.5, 2), (0, 0.3))),
                       ((np.array([-1., 0., 0.5, 3.]), 5, 2), None)):
        edges = stat_tools._edges(x, bins, rng_)
        ref = np.histogramdd(data, bins, rng_)[1]
        assert len(edges) == 3 and all(np.array_equal(e, r) for e, r in zip(edges, ref))
    flat = torch.from_numpy(np.full((5, 2), 0.25))
    assert all(np.array_equal(e, r) for e, r in zip(stat_tools._edges(flat, 4, None),
                                                   np.histogramdd(np.full((5, 2), 0.25), 4)[1]))
    with pytest.raises(ValueError):
        stat_tools._edges(x, (3, 4), None)
    with pytest.raises(ValueError):
        stat_tools._edges(x, 0, None)
    srt = torch.sort(x, dim=0).values
    for q in (0.25, 0.5, 0.75, 0.0, 1.0, 0.333):
        for d in range(3):
            assert stat_tools._percentile(srt[:, d], q) == np.percentile(data[:, d], 100 * q)


def test_interfaces_namespace():
    """Recipe names and signatures of hepmc/interfaces/*.py; NUTS recipes are declared out of scope."""
    import inspect
    import hepmc_b200 as H
    names = ["plain_uniform", "plain_hmc", "plain_nuts", "plain_spherical_nuts", "mc3_gauss", "mc3_hmc",
             "mc3_nuts", "mc3_spherical_nuts", "mc3_hmc_el", "mc3_nuts_el", "mc3_spherical_nuts_el", "mc3_hmc_uni"]
    assert sorted(names) == H.interfaces.__all__
    sig = inspect.signature(H.interfaces.mc3_hmc_el.get_sampler)
    assert list(sig.parameters) == ["target", "ndim", "initial", "centers", "widths", "beta", "nintegral", "mass",
                                    "steps", "step_size", "el_size"]
    assert sig.parameters["el_size"].default == 100 and sig.parameters["mass"].default == 10
    assert list(inspect.signature(H.interfaces.mc3_gauss.get_sampler).parameters)[-1] == "local_width"
    with pytest.raises(NotImplementedError):
        H.interfaces.plain_nuts.get_sampler(None, 2, 0.5)


def test_sample_summaries_single_chain_and_ensemble():
    """Sample.size / ndim / mean / variance: the reference's meaning for (T, ndim) data; an ensemble
    (C, T, ndim) reports the chain length and pools the states."""
    import hepmc_b200 as H
    rs = np.random.RandomState(0)
    one = H.core.sampling.Sample(data=rs.rand(50, 3))
    assert (one.size, one.ndim) == (50, 3)
    assert np.allclose(one.mean, one.data.mean(axis=0)) and np.allclose(one.variance, one.data.var(axis=0))
    ens = H.core.sampling.Sample(data=rs.rand(7, 50, 3))
    assert (ens.size, ens.ndim) == (50, 3)
    assert np.allclose(ens.mean, ens.data.reshape(-1, 3).mean(axis=0))
    assert np.allclose(ens.variance, ens.data.reshape(-1, 3).var(axis=0))
    assert H.core.sampling.Sample().size is None and "N/A" in repr(H.core.sampling.Sample())


def test_vegas_argument_validation_is_at_construction():
    import hepmc_b200 as H
    with pytest.raises(ValueError):
        H.VegasMC(2, divisions=256)
    with pytest.raises(ValueError):
        H.VegasMC(33, divisions=4)
    H.VegasMC(32, divisions=255)


def test_rng_modes_host_state():
    import hepmc_b200 as H
    assert H.rng.get_mode() == H.rng.DEFAULT_MODE == "u32r10" and H.rng.mode_code() == 1
    H.rng.set_mode("u52r10")
    assert H.rng.mode_code() == 0
    H.seed(5)                                   # the mode is a setting, not part of the seed state
    assert H.rng.get_mode() == "u52r10"
    with pytest.raises(ValueError):
        H.rng.set_mode("u16r1")
    H.rng.set_mode()
    assert H.rng.get_mode() == "u32r10"
