"""Full-size runs of the BASELINE.json configurations, checked through size-independent
properties (the oracle cannot run these sizes): determinism, known integrals and moments,
momentum conservation, agreement between precisions.  A few seconds on one B200."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import hepmc_b200 as H


def test_c3_vegas_camel16_full_size():
    """camel-16D VEGAS, divisions 15, 1e9 points per iteration: same seed -> bit-identical result;
    the estimate agrees with the analytic mass of the camel inside the unit cube."""
    n, iters = 1_000_000_000, 4
    out = []
    for _ in range(2):
        H.seed(2024)
        mc = H.VegasMC(16, divisions=15, c=3)
        out.append(mc(H.densities.Camel(16), n, iters, keep_samples=False))
    assert out[0].integral == out[1].integral and out[0].integral_err == out[1].integral_err
    sigma = math.sqrt(0.005)
    inside_a = 0.5 * (math.erf((1 - 1 / 3) / sigma / math.sqrt(2)) + math.erf((1 / 3) / sigma / math.sqrt(2)))
    truth = inside_a ** 16       # both modes carry the same mass by symmetry
    assert abs(out[0].integral - truth) < 4 * out[0].integral_err
    # unweighted mean over the iterations (vegas.py:133-136): the first ones, on the still uniform
    # grid of a 16-D double peak, dominate the error
    assert out[0].integral_err < 0.3
    # the grid starts to concentrate on the region of the modes
    sizes = np.asarray(mc.volumes.sizes)
    bounds = np.asarray(mc.volumes.bounds)
    k = np.argmin(sizes, axis=1)
    centres = 0.5 * (bounds[np.arange(16), k] + bounds[np.arange(16), k + 1])
    assert np.all((centres > 0.2) & (centres < 0.8))       # damped (c = 3): 4 iterations only start to adapt
    assert np.all(sizes.min(axis=1) < 0.95 / 15) and np.allclose(sizes.sum(axis=1), 1.0, atol=1e-12)


def test_c2_rambo_full_launch_conserves_momentum():
    """RAMBO 2 -> 4 at sqrt(s) = 100: 2^27 points in one launch (16 GiB); every point conserves
    four-momentum and is massless; a second launch with the same counters is bit-identical."""
    H.set_outputs("torch")
    try:
        m = H.phase_space.Rambo(100., 4)
        n = 1 << 27
        H.seed(5)
        p = m.generate(n, stream_id=3)
        P = p.view(n, 4, 4)
        tot = P.sum(dim=1)
        assert float((tot[:, 0] - 100.).abs().max()) < 1e-10
        assert float(tot[:, 1:].abs().max()) < 1e-10
        m2 = P[:, :, 0] ** 2 - (P[:, :, 1:] ** 2).sum(dim=2)
        assert float(m2.abs().max()) < 1e-8 * 100. ** 2
        assert bool((P[:, :, 0] > 0).all())
        check = p[:: 1 << 12].clone()
        del p, P, tot, m2
        q = m.generate(n, stream_id=3)
        assert torch.equal(q[:: 1 << 12], check)
        # a slice generated on its own (another "rank") equals the same rows of the whole
        part = m.generate(1 << 20, first_index=5 << 20, stream_id=3)
        assert torch.equal(part, q[5 << 20: 6 << 20])
    finally:
        H.set_outputs("numpy")


def test_c4_rwm_banana_full_size():
    """banana-2D RWM, 1e6 chains x 1e4 steps (cov = 10 I, start (0, 10)): ensemble moments of the
    final states are the banana's (x0 ~ N(0, 100), x1 = z - b (x0^2 - 100), z ~ N(0, 1))."""
    H.set_outputs("torch")
    try:
        H.seed(11)
        C, T = 1_000_000, 10_001
        met = H.DefaultMetropolis(2, H.densities.Banana(2), H.proposals.Gaussian(2, cov=10 * np.eye(2)))
        x0 = torch.tensor([[0., 10.]], device="cuda", dtype=torch.float64).repeat(C, 1)
        s = met.sample(T, x0, store_chain=False)
        x = s.data[:, 0]
        rate = float(s.accepted.double().mean()) / T
        assert 0.15 < rate < 0.30
        assert int(s.total_accepted.item()) == int(s.accepted.sum())
        mean, var = x.mean(dim=0).cpu().numpy(), x.var(dim=0).cpu().numpy()
        assert abs(mean[0]) < 0.1 and abs(var[0] / 100. - 1) < 0.02
        b = 0.1
        # E[x1] = -b (E[x0^2] - 100) = 0 ; Var[x1] = 1 + b^2 Var[x0^2] = 1 + b^2 * 2 * 100^2
        assert abs(mean[1]) < 0.1 and abs(var[1] / (1 + b * b * 2e4) - 1) < 0.03
    finally:
        H.set_outputs("numpy")


def test_c5_hmc_elm_full_size_precisions_agree():
    """camel-8D HMC with a trained 1024-node Gaussian-basis surrogate gradient, 1e5 chains x 20
    leapfrog: float64, float32 and tensor-core gradients give the same acceptance rate (the
    Metropolis test uses the float64 pdf in all three) and the same ensemble moments -- for a
    surrogate whose gradient sum does not cancel by more than ~1e3 (ridge-regularised training).
    The reference's plain pseudo-inverse at this node count cancels by ~1e4: float32 and the
    "tc32" mode (FP32 contraction) still agree; "tc" detects the cancellation, warns and runs tc32."""
    import warnings
    H.set_outputs("torch")
    try:
        rs = np.random.default_rng(0)
        nodes = 1024
        basis = H.surrogate.GaussianBasis(8)
        tgt0 = H.densities.UnconstrainedCamel(8)
        xs = np.concatenate([1 / 3 + 0.07 * rs.standard_normal((3000, 8)), 2 / 3 + 0.07 * rs.standard_normal((3000, 8))])
        hidden = (rs.random((nodes, 8)), 0.01 + 0.99 * rs.random(nodes), None)   # the reference's ranges
        C, T = 100_000, 6
        x0 = torch.tensor(np.where(rs.random((C, 1)) < 0.5, 1 / 3, 2 / 3) + 0.05 * rs.standard_normal((C, 8)), device="cuda")

        def run(ridge):
            params, bias, w = basis.extreme_learning_train(xs, tgt0.pot(xs), nodes, params=hidden, ridge=ridge)
            rates, means, kappa, warned = {}, {}, None, False
            for prec in ("f64", "f32", "tc32", "tc"):
                H.seed(77)
                tgt = H.densities.Camel(8)
                grad = H.surrogate.attach_surrogate_gradient(tgt, basis, params, bias, w)
                kappa = grad.cancellation(x0[:256])
                upd = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8), 20, 0.01, precision=prec)
                with warnings.catch_warnings(record=True) as caught:
                    warnings.simplefilter("always")
                    s = upd.sample(T, x0, store_chain=False)
                warned = warned or any(issubclass(c.category, RuntimeWarning) and "cancels" in str(c.message) for c in caught)
                rates[prec] = float(s.accepted.double().mean()) / T
                means[prec] = s.data[:, 0].mean(dim=0).cpu().numpy()
            return rates, means, kappa, warned

        rates, means, kappa, warned = run(1e-6)
        assert kappa < 500 and not warned, kappa
        assert rates["f64"] > 0.6, rates
        assert abs(rates["f32"] - rates["f64"]) < 0.005 and abs(rates["tc"] - rates["f64"]) < 0.01, rates
        assert np.allclose(means["f32"], means["f64"], atol=2e-3) and np.allclose(means["tc"], means["f64"], atol=2e-3)
        assert abs(rates["tc32"] - rates["f64"]) < 0.005 and np.allclose(means["tc32"], means["f64"], atol=2e-3)
        # plain pseudo-inverse: "tc" notices the cancellation, warns and runs the tc32 kernel
        rates, means, kappa, warned = run(0.0)
        assert kappa > H.hamiltonian.hmc.TC_CANCELLATION_LIMIT and warned, kappa
        assert abs(rates["f32"] - rates["f64"]) < 0.005, rates
        assert abs(rates["tc32"] - rates["f64"]) < 0.01 and rates["tc"] == rates["tc32"], rates
    finally:
        H.set_outputs("numpy")
