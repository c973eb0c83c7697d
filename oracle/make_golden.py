"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (authoring container only).

    python oracle/make_golden.py

Imports hepmc from /root/reference/src through oracle/refshim.py, records the
random draws each call consumes (np.random.* wrapped by recorders that forward
to the real legacy generator) and stores inputs + reference outputs as small
fixtures.  The fixtures pin the NumPy oracle (tests/test_oracle_golden.py) and,
through it and directly, the CUDA kernels (tests/test_gpu_*.py, replay mode).

Nothing here is product code; the GPU box never runs this file.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED = 1234  # repository convention (examples/*.py)


class Tape(object):
    """Wrap np.random functions so the draws a reference call consumes are logged."""

    def __init__(self):
        self.log = {}
        self._saved = {}

    def _wrap(self, name, fn=None):
        orig = getattr(np.random, name)
        self._saved[name] = orig
        log = self.log.setdefault(name, [])

        def wrapped(*a, **k):
            out = (fn or orig)(*a, **k)
            log.append(np.array(out, copy=True))
            return out
        setattr(np.random, name, wrapped)

    def __enter__(self):
        for name in ("random", "rand", "randint"):
            self._wrap(name)
        # legacy multivariate_normal == mean + sqrt(diag cov) * standard_normal for
        # diagonal covariances (asserted in check_mvn below).
        self.log["z"] = []
        orig_mvn = np.random.multivariate_normal
        self._saved["multivariate_normal"] = orig_mvn
        zlog = self.log["z"]

        def mvn(mean, cov, size=None):
            mean = np.asarray(mean, dtype=np.float64)
            cov = np.asarray(cov, dtype=np.float64)
            assert np.count_nonzero(cov - np.diag(np.diagonal(cov))) == 0
            scale = np.sqrt(np.diagonal(cov))
            if size is None:
                z = np.random.standard_normal(mean.shape[0])
                zlog.append(z.copy())
                return mean + scale * z
            z = np.random.standard_normal((size, mean.shape[0]))
            zlog.append(z.copy())
            return mean + scale * z
        np.random.multivariate_normal = mvn
        return self

    def __exit__(self, *exc):
        for name, fn in self._saved.items():
            setattr(np.random, name, fn)


def check_mvn():
    """App. A.7: the patched proposal equals NumPy's legacy multivariate_normal."""
    for mean, cov in (([0.3, -2.], 10 * np.eye(2)), (np.zeros(8), np.eye(8)),
                      ([1., 2.], np.diag([100., 1.]))):
        np.random.seed(7)
        a = np.random.multivariate_normal(mean, cov)
        a1 = np.random.multivariate_normal(mean, cov, 1)
        np.random.seed(7)
        z = np.random.standard_normal(len(mean))
        z1 = np.random.standard_normal((1, len(mean)))
        b = np.asarray(mean) + np.sqrt(np.diagonal(cov)) * z
        b1 = np.asarray(mean) + np.sqrt(np.diagonal(cov)) * z1
        assert np.array_equal(a, b) and np.array_equal(a1, b1), "legacy MVN mismatch"


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def points(n, ndim, rng, spread=True):
    """Test points: inside the cube, near the modes, and some outside / on the edge."""
    xs = rng.random((n, ndim))
    k = n // 4
    xs[:k] = 1 / 3 + 0.08 * rng.standard_normal((k, ndim))
    xs[k:2 * k] = 2 / 3 + 0.08 * rng.standard_normal((k, ndim))
    if spread:
        xs[-6:] = rng.uniform(-0.5, 1.5, (6, ndim))
        xs[-7, 0] = 0.0
        xs[-8, -1] = 1.0
    return xs


def gen_densities(h):
    rng = np.random.default_rng(SEED)
    out = {}
    D = h.densities
    for ndim in (1, 2, 3, 8, 16):
        xs = points(64, ndim, rng)
        out["xs_%d" % ndim] = xs
        for cls, tag in ((D.Camel, "camel"), (D.UnconstrainedCamel, "ucamel")):
            d = cls(ndim)
            out["%s_pdf_%d" % (tag, ndim)] = d.pdf(xs)
            out["%s_pdfgrad_%d" % (tag, ndim)] = d.pdf_gradient(xs)
            out["%s_pot_%d" % (tag, ndim)] = d.pot(xs)
            with np.errstate(all='ignore'):
                out["%s_potgrad_%d" % (tag, ndim)] = d.pot_gradient(xs)
        g = D.Gaussian(ndim, mu=0.25, cov=0.7)
        out["gauss_pdf_%d" % ndim] = g.pdf(xs)
        out["gauss_pot_%d" % ndim] = g.pot(xs)
        out["gauss_potgrad_%d" % ndim] = g.pot_gradient(xs)
        cov = np.linspace(0.5, 2.0, ndim)
        mu = np.linspace(-0.2, 0.4, ndim)
        g = D.Gaussian(ndim, mu=mu, cov=cov)
        out["gaussd_cov_%d" % ndim] = cov
        out["gaussd_mu_%d" % ndim] = mu
        out["gaussd_pdf_%d" % ndim] = g.pdf(xs)
        out["gaussd_pot_%d" % ndim] = g.pot(xs)
        out["gaussd_potgrad_%d" % ndim] = g.pot_gradient(xs)
        u = D.Uniform(ndim)
        out["uniform_pdf_%d" % ndim] = u.pdf(xs).astype(np.float64)
        out["uniform_pot_%d" % ndim] = np.asarray(u.pot(xs), dtype=np.float64)
    # far tails (maha ~ 700) -- the exp-argument parity case of SURVEY.md 7.2
    xs = np.full((4, 16), 1 / 3) + np.array([[0.6], [0.62], [0.64], [0.66]])
    out["xs_tail"] = xs
    out["ucamel_pdf_tail"] = D.UnconstrainedCamel(16).pdf(xs)
    # banana
    for ndim in (2, 3):
        xs = np.column_stack([rng.normal(0, 10, 64), rng.normal(5, 6, 64)] +
                             [rng.normal(0, 1, 64) for _ in range(ndim - 2)])
        b = D.Banana(ndim)
        out["banana_xs_%d" % ndim] = xs
        out["banana_pdf_%d" % ndim] = b.pdf(xs)
        out["banana_pot_%d" % ndim] = b.pot(xs)
    out["banana_potgrad_2"] = D.Banana(2).pot_gradient(out["banana_xs_2"])
    # Density.__call__ adaptor (density.py:39-46)
    xs = out["xs_3"]
    out["camel_call_3"] = D.Camel(3)(*xs.transpose())
    save("densities", **out)


def gen_plain(h):
    out = {}
    for ndim, n in ((2, 4096), (16, 1024)):
        np.random.seed(SEED)
        with Tape() as t:
            s = h.PlainMC(ndim)(h.densities.Camel(ndim), n)
        out["data_%d" % ndim] = t.log["random"][0]
        assert np.array_equal(s.data, t.log["random"][0])
        out["f_%d" % ndim] = s.function_values
        out["integral_%d" % ndim] = s.integral
        out["err_%d" % ndim] = s.integral_err
    save("plain_mc", **out)


def gen_vegas(h):
    from hepmc.core.integration import stratified_volume as sv
    out = {}
    cases = (("2d", 2, 50, 3, 2000, 5, False), ("16d", 16, 15, 3, 1000, 3, False),
             ("3d_vw", 3, 8, 1.5, 1500, 4, True))
    for tag, ndim, K, c, n, iters, vw in cases:
        np.random.seed(SEED)
        mc = h.VegasMC(ndim, divisions=K, c=c, var_weighted=vw)
        hist = [np.stack(mc.volumes.bounds).copy()]
        orig = sv.GridVolumes.sizes.fset

        def logged(self, sizes, _orig=orig, _hist=hist):
            # App. B2: bounds[d][0] is np.empty garbage in the reference; force 0
            _orig(self, sizes)
            for d in range(self.ndim):
                self.bounds[d][0] = 0.0
            _hist.append(np.stack(self.bounds).copy())
        sv.GridVolumes.sizes = property(sv.GridVolumes.sizes.fget, logged)
        try:
            with Tape() as t:
                s = mc(h.densities.Camel(ndim), n, iters, chi=True)
        finally:
            sv.GridVolumes.sizes = property(sv.GridVolumes.sizes.fget, orig)
        idx = np.stack(t.log["randint"]).reshape(iters, ndim, n)
        u = np.stack(t.log["rand"]).reshape(iters, ndim, n)
        out["idx_" + tag] = idx.astype(np.int64)
        out["u_" + tag] = u
        out["bounds_" + tag] = np.stack(hist)
        out["integral_" + tag] = s.integral
        out["err_" + tag] = s.integral_err
        out["chi2_" + tag] = s.chi2
        out["data_" + tag] = s.data
        out["f_" + tag] = s.function_values
        out["w_" + tag] = s.weights
        out["params_" + tag] = np.array([ndim, K, c, n, iters, int(vw)], dtype=np.float64)
    # GridVolumes.pdf on an adapted grid + ImportanceMC(GridVolumes)
    np.random.seed(SEED)
    mc = h.VegasMC(2, divisions=20)
    mc(h.densities.Camel(2), 2000, 4)
    for d in range(2):
        mc.volumes.bounds[d][0] = 0.0
    grid = mc.volumes
    out["gridpdf_bounds"] = np.stack(grid.bounds)
    rng = np.random.default_rng(SEED)
    xs = points(64, 2, rng)
    out["gridpdf_xs"] = xs
    out["gridpdf_pdf"] = grid.pdf(xs)
    with Tape() as t:
        s = h.ImportanceMC(grid)(h.densities.Camel(2), 3000)
    out["imp_idx"] = np.stack(t.log["randint"]).astype(np.int64)
    out["imp_u"] = np.stack(t.log["rand"])
    out["imp_data"] = s.data
    out["imp_f"] = s.function_values
    out["imp_w"] = s.weights
    out["imp_integral"] = s.integral
    out["imp_err"] = s.integral_err
    save("vegas", **out)


def gen_rambo(h):
    out = {}
    rng = np.random.default_rng(SEED)
    for tag, e_cm, npart, n in (("4_100", 100., 4, 512), ("3_50", 50., 3, 128),
                                ("6_1000", 1000., 6, 128), ("2_10", 10., 2, 64)):
        xs = rng.random((n, 4 * npart))
        m = h.phase_space.Rambo(e_cm, npart)
        out["xs_" + tag] = xs
        out["p_" + tag] = m.map(xs)
        out["pdf_" + tag] = m.pdf(xs)
    # densities.Rambo.rvs + ImportanceMC as in examples/rambo.py:21-27
    np.random.seed(SEED)
    dist = h.densities.Rambo(4, 100.)
    with Tape() as t:
        s = h.ImportanceMC(dist)(lambda *p: p[0] * p[4] / 1e3 + p[3] ** 2 / 1e3, 2048)
    out["imp_xs"] = t.log["random"][0]
    out["imp_data"] = s.data
    out["imp_f"] = s.function_values
    out["imp_w"] = np.float64(s.weights)
    out["imp_integral"] = s.integral
    out["imp_err"] = s.integral_err
    save("rambo", **out)


def run_chains(make_update, x0s, T):
    """Run the reference sampler once per start; returns chain, accepted, z tape, u tape.

    u is recorded sparsely (drawn only if accept<1): the tape stores NaN where
    no uniform was consumed, so a replay that always reads u[t] must never need it.
    """
    chains, accs, zs, us = [], [], [], []
    for x0 in x0s:
        upd = make_update()
        with Tape() as t:
            # count uniforms per step by wrapping next_state
            marks = []
            orig_next = upd.next_state

            def next_state(state, it, _o=orig_next, _m=marks, _t=t):
                before = len(_t.log["rand"])
                out = _o(state, it)
                _m.append(len(_t.log["rand"]) - before)
                return out
            upd.next_state = next_state
            s = upd.sample(T, x0, log_every=-1)
        chains.append(s.data)
        accs.append(s.accepted)
        z = np.stack([np.ravel(v) for v in t.log["z"]])
        zs.append(z)
        u = np.full(T - 1, np.nan)
        flat = [float(v) for v in t.log["rand"]]
        k = 0
        for i, m in enumerate(marks):
            assert m in (0, 1)
            if m:
                u[i] = flat[k]
                k += 1
        us.append(u)
    return np.stack(chains), np.array(accs, dtype=np.int64), zs, np.stack(us)


def gen_metropolis(h):
    out = {}
    np.random.seed(SEED)
    T = 400
    x0s = [np.array([0., 10.]), np.array([-20., -30.]), np.array([3., 7.]), np.array([-5., 9.])]

    def mk():
        return h.DefaultMetropolis(2, h.densities.Banana(2),
                                   h.proposals.Gaussian(2, cov=10 * np.eye(2)))
    chain, acc, zs, u = run_chains(mk, x0s, T)
    out["banana_x0"] = np.stack(x0s)
    out["banana_chain"] = chain
    out["banana_accepted"] = acc
    out["banana_z"] = np.stack(zs)            # (C, T-1, 2)
    out["banana_u"] = u
    out["banana_cov"] = np.array([10., 10.])
    # camel-2D target, Gaussian proposal, anisotropic diagonal covariance
    x0s = [np.array([0.3, 0.35]), np.array([0.7, 0.6])]

    def mk2():
        return h.DefaultMetropolis(2, h.densities.Camel(2),
                                   h.proposals.Gaussian(2, cov=np.diag([0.01, 0.02])))
    chain, acc, zs, u = run_chains(mk2, x0s, T)
    out["camel_x0"] = np.stack(x0s)
    out["camel_chain"] = chain
    out["camel_accepted"] = acc
    out["camel_z"] = np.stack(zs)
    out["camel_u"] = u
    out["camel_cov"] = np.array([0.01, 0.02])
    # gaussian-3D target
    x0s = [np.array([0.1, -0.2, 0.3])]

    def mk3():
        return h.DefaultMetropolis(3, h.densities.Gaussian(3, mu=0.25, cov=0.7),
                                   h.proposals.Gaussian(3, cov=0.5 * np.eye(3)))
    chain, acc, zs, u = run_chains(mk3, x0s, T)
    out["gauss_x0"] = np.stack(x0s)
    out["gauss_chain"] = chain
    out["gauss_accepted"] = acc
    out["gauss_z"] = np.stack(zs)
    out["gauss_u"] = u
    out["gauss_cov"] = np.array([0.5, 0.5, 0.5])
    # camel-2D target, UniformLocal proposal (proposals/uniform_local.py): per step the proposal's
    # scalar uniform, then (if the ratio is below one) the accept uniform
    x0s = [np.array([0.3, 0.35]), np.array([0.95, 0.02])]
    delta = np.array([0.2, 0.3])
    chains, accs, vs, us = [], [], [], []
    for x0 in x0s:
        upd = h.DefaultMetropolis(2, h.densities.Camel(2), h.proposals.UniformLocal(2, delta))
        with Tape() as t:
            marks = []
            orig_next = upd.next_state

            def next_state(state, it, _o=orig_next, _m=marks, _t=t):
                before = len(_t.log["rand"])
                out_state = _o(state, it)
                _m.append(len(_t.log["rand"]) - before)
                return out_state
            upd.next_state = next_state
            smp = upd.sample(T, x0, log_every=-1)
        flat = [float(v) for v in t.log["rand"]]
        v, u, k = np.empty(T - 1), np.full(T - 1, np.nan), 0
        for i, m in enumerate(marks):
            assert m in (1, 2)
            v[i] = flat[k]
            if m == 2:
                u[i] = flat[k + 1]
            k += m
        chains.append(smp.data); accs.append(smp.accepted); vs.append(v); us.append(u)
    out["ulocal_x0"] = np.stack(x0s)
    out["ulocal_chain"] = np.stack(chains)
    out["ulocal_accepted"] = np.array(accs, dtype=np.int64)
    out["ulocal_v"] = np.stack(vs)
    out["ulocal_u"] = np.stack(us)
    out["ulocal_delta"] = delta
    save("metropolis", **out)


def gen_elm_and_hmc(h):
    from hepmc.surrogate import extreme_learning as el
    out = {}
    np.random.seed(SEED)
    ndim = 8
    target = h.densities.Camel(ndim)
    utarget = h.densities.UnconstrainedCamel(ndim)
    # training set as in SURVEY.md 8(d) C5 (scaled down): points around both modes
    n_train = 600
    xs = np.concatenate([1 / 3 + np.sqrt(0.005) * np.random.standard_normal((n_train // 2, ndim)),
                         2 / 3 + np.sqrt(0.005) * np.random.standard_normal((n_train // 2, ndim))])
    values = utarget.pot(xs)
    for tag, nodes in (("g64", 64), ("g100", 100)):
        basis = el.GaussianBasis(ndim)
        params, bias, weights = basis.extreme_learning_train(xs, values, nodes)
        centers, widths, _ = params
        q = points(48, ndim, np.random.default_rng(SEED + 1), spread=False)
        out["%s_centers" % tag] = centers
        out["%s_widths" % tag] = widths
        out["%s_bias" % tag] = bias
        out["%s_weights" % tag] = weights
        out["%s_q" % tag] = q
        out["%s_eval" % tag] = basis.eval(params, bias, weights, q)
        out["%s_grad" % tag] = basis.eval_gradient(params, bias, weights, q)
        out["%s_outmat" % tag] = basis.output_matrix(q, params)
    out["train_xs"] = xs
    out["train_values"] = values
    tb = el.TrigBasis(ndim)
    tparams, tbias, tweights = tb.extreme_learning_train(xs, values, 48)
    q = out["g64_q"]
    out["t48_biases"] = tparams[0]
    out["t48_inweights"] = tparams[1]
    out["t48_bias"] = tbias
    out["t48_weights"] = tweights
    out["t48_eval"] = tb.eval(tparams, tbias, tweights, q)
    out["t48_grad"] = tb.eval_gradient(tparams, tbias, tweights, q)
    save("elm", **out)

    # ---- HMC ----
    hout = {}
    T = 60

    def chains_for(make_update, x0s):
        chain, acc, zs, u = run_chains(make_update, x0s, T)
        # z tape: first entry is the init_state momentum (hmc.py:86-87), then one per step
        p = np.stack([z[1:] for z in zs])
        return chain, acc, p, u

    # (a) exact gradient, unit mass (examples/test_hmc.py:20-24 parameters)
    x0s = [np.full(ndim, 1 / 3), np.full(ndim, 2 / 3), np.linspace(0.3, 0.4, ndim)]

    def mk_exact():
        return h.hamiltonian.HamiltonianUpdate(
            h.densities.Camel(ndim), h.densities.Gaussian(ndim), 20, 0.02)
    chain, acc, p, u = chains_for(mk_exact, x0s)
    hout["exact_x0"] = np.stack(x0s)
    hout["exact_chain"] = chain
    hout["exact_accepted"] = acc
    hout["exact_p"] = p           # N(0, I) momenta (unit mass: p == z)
    hout["exact_u"] = u
    hout["exact_params"] = np.array([20, 0.02, 1.0])
    # (b) ELM surrogate gradient, mass 10 (interfaces/mc3_hmc_el.py:9,32-40 wiring)
    centers, widths = out["g100_centers"], out["g100_widths"]
    bias, weights = out["g100_bias"], out["g100_weights"]
    basis = el.GaussianBasis(ndim)

    def mk_elm():
        tgt = h.densities.Camel(ndim)
        tgt.pot_gradient = lambda xs: basis.eval_gradient(
            (centers, widths, None), bias, weights, xs)
        return h.hamiltonian.HamiltonianUpdate(
            tgt, h.densities.Gaussian(ndim, cov=10.), 10, 0.1)
    chain, acc, p, u = chains_for(mk_elm, x0s)
    hout["elm_x0"] = np.stack(x0s)
    hout["elm_chain"] = chain
    hout["elm_accepted"] = acc
    hout["elm_p"] = p * np.sqrt(10.)      # tape stores the momenta themselves
    hout["elm_u"] = u
    hout["elm_params"] = np.array([10, 0.1, 10.0])
    # (c) banana 2-D exact gradient (examples/hmc_banana.py)
    x0s = [np.array([0., 10.]), np.array([2., 6.])]

    def mk_banana():
        return h.hamiltonian.HamiltonianUpdate(
            h.densities.Banana(2), h.densities.Gaussian(2), 10, 0.1)
    chain, acc, p, u = chains_for(mk_banana, x0s)
    hout["banana_x0"] = np.stack(x0s)
    hout["banana_chain"] = chain
    hout["banana_accepted"] = acc
    hout["banana_p"] = p
    hout["banana_u"] = u
    hout["banana_params"] = np.array([10, 0.1, 1.0])
    save("hmc", **hout)


def gen_leapfrog(h):
    """HamiltonLeapfrog.__call__ (hamiltonian/simulation.py:26-46) on its own: one state per call, as
    the reference uses it; exact camel gradient, ELM surrogate gradient (the weights of elm.npz),
    banana.  Inputs and outputs only -- the call draws nothing."""
    from hepmc.hamiltonian.simulation import HamiltonLeapfrog
    from hepmc.surrogate import extreme_learning as el
    out = {}
    rs = np.random.RandomState(SEED + 7)
    ndim = 8
    elm = dict(np.load(os.path.join(OUT, "elm.npz")))

    def run(tag, pot_gradient, kin_gradient, eps, steps, q0s, p0s):
        sim = HamiltonLeapfrog(pot_gradient, kin_gradient, eps, steps)
        qs, ps = [], []
        for q0, p0 in zip(q0s, p0s):
            q, p = sim(q0, p0)
            qs.append(np.array(q, dtype=np.float64))
            ps.append(np.array(p, dtype=np.float64))
        out[tag + "_q0"], out[tag + "_p0"] = np.stack(q0s), np.stack(p0s)
        out[tag + "_q"], out[tag + "_p"] = np.stack(qs), np.stack(ps)
        out[tag + "_params"] = np.array([steps, eps])

    n = 24
    q0s = np.where(rs.randint(0, 2, (n, 1)) == 0, 1 / 3, 2 / 3) + 0.06 * rs.standard_normal((n, ndim))
    # (a) exact camel gradient, unit mass: the pair HamiltonianUpdate builds (hmc.py:47-50)
    camel = h.densities.Camel(ndim)
    run("exact", camel.pot_gradient, h.densities.Gaussian(ndim).pot_gradient, 0.02, 20, q0s, rs.standard_normal((n, ndim)))
    out["exact_mass"] = np.ones(ndim)
    # (b) ELM surrogate gradient, mass 10 (interfaces/mc3_hmc_el.py:32-40 wiring)
    basis = el.GaussianBasis(ndim)
    grad = lambda xs: basis.eval_gradient((elm["g100_centers"], elm["g100_widths"], None), elm["g100_bias"],
                                          elm["g100_weights"], xs)
    run("elm", grad, h.densities.Gaussian(ndim, cov=10.).pot_gradient, 0.1, 10, q0s,
        np.sqrt(10.) * rs.standard_normal((n, ndim)))
    out["elm_mass"] = np.full(ndim, 10.)
    # (c) banana 2-D (examples/hmc_banana.py), zero steps included (identity)
    qb = np.stack([[0., 10.], [2., 6.], [-3., 4.], [0.5, 9.5]])
    run("banana", h.densities.Banana(2).pot_gradient, h.densities.Gaussian(2).pot_gradient, 0.1, 10, qb,
        rs.standard_normal((4, 2)))
    out["banana_mass"] = np.ones(2)
    run("banana0", h.densities.Banana(2).pot_gradient, h.densities.Gaussian(2).pot_gradient, 0.1, 0, qb,
        rs.standard_normal((4, 2)))
    save("leapfrog", **out)


def gen_ks_reference():
    """Regenerate testsuit/reference_camel-1M.2d.npy (missing blob) as its 15x15
    histogram + edges -- what testsuit/testing/Analyse.py:37-38 makes of it --
    and validate against the KS values stored in testsuit/Camel2d-cv.npy."""
    rng = np.random.default_rng(SEED)
    need = 1_000_000
    pts = np.empty((0, 2))
    while pts.shape[0] < need:
        m = 1_200_000
        comp = rng.integers(0, 2, m)
        x = np.where(comp[:, None] == 0, 1 / 3, 2 / 3) + np.sqrt(0.005) * rng.standard_normal((m, 2))
        x = x[np.all((x > 0) & (x < 1), axis=1)]
        pts = np.concatenate([pts, x])
    pts = pts[:need]
    truth, edges = np.histogramdd(pts, bins=15)
    out = dict(truth=truth, edges=np.stack(edges), mean=pts.mean(axis=0), var=pts.var(axis=0))
    # validation against stored analysis
    sys.path.insert(0, "/root/reference/testsuit")
    from testing.stats import ks
    checks = []
    for fam, beta in (("cv", "0.4"), ("hmc", "0.8")):
        name = "Camel_100.0k_%s-beta_%s.npy" % ("vanila" if fam == "cv" else "hmc", beta)
        smp = np.load("/root/reference/testsuit/%s/%s" % (fam, name), allow_pickle=True)[()]["samples"]
        stored = np.load("/root/reference/testsuit/Camel2d-%s.npy" % fam, allow_pickle=True)
        row = [r for r in stored if r["Sample"].endswith(name) and r["lag"] == 1][0]
        size = int(round(len(smp) / 20))
        hist, _ = np.histogramdd(smp[::1][:size], bins=edges)
        mine = np.array(ks(hist, truth.copy()))
        theirs = np.array(row["ks"])
        print("KS check %s beta=%s: regenerated %s stored %s" % (fam, beta, mine[:, 0], theirs[:, 0]))
        # the bin edges come from the min/max of the (re)generated 1M sample, so
        # d_max moves by O(bin shift * marginal density) ~ 5e-3 between two
        # statistically equivalent reference samples; d_alpha(n=5000) = 0.0192.
        assert np.all(np.abs(mine[:, 0] - theirs[:, 0]) < 1e-2)
        checks.append(np.concatenate([mine[:, 0], theirs[:, 0]]))
        # keep the binned sample so the oracle's ks() can be pinned without the blob
        out["chain_hist_%s" % fam] = hist
        out["chain_ks_%s" % fam] = mine
    out["ks_checks"] = np.stack(checks)
    save("camel2d_reference_hist", **out)


def gen_multichannel(h):
    """MultiChannelMC (importance.py:54-230) with two Gaussian channels and a uniform one."""
    from hepmc.core.integration import multi_channel as mcm
    out = {}
    np.random.seed(SEED)
    D = h.densities
    chans = [D.Gaussian(2, mu=1 / 3, scale=0.1), D.Gaussian(2, mu=[0.6, 0.7], scale=[0.15, 0.1]), D.Uniform(2)]
    mc = h.MultiChannel(chans)
    tapes, alphas = [], []
    orig = mcm.MultiChannel.rvs

    def rvs(self, sample_size, return_sizes=False):
        pts, sizes = orig(self, sample_size, True)
        tapes.append((pts.copy(), sizes.copy()))
        alphas.append(self.channels_weight.copy())
        return (pts, sizes) if return_sizes else pts
    mcm.MultiChannel.rvs = rvs
    try:
        s = h.MultiChannelMC(mc, b=0.5)(D.Camel(2), [400, 400], [500, 500, 500], [600])
    finally:
        mcm.MultiChannel.rvs = orig
    for j, (pts, sizes) in enumerate(tapes):
        out["xs_%d" % j] = pts
        out["sizes_%d" % j] = sizes.astype(np.int64)
        out["alpha_%d" % j] = alphas[j]
    out["alpha_final"] = mc.channels_weight.copy()
    out["n_iter"] = np.array([2, 3, 1])
    out["integral"] = s.integral
    out["err"] = s.integral_err
    out["data"] = s.data
    out["f"] = s.function_values
    out["w"] = s.weights
    out["chan_mu1"] = np.array([0.6, 0.7])
    out["chan_cov1"] = np.array([0.15, 0.1]) ** 2
    # MultiChannel.pdf on fixed points with fixed weights
    rng = np.random.default_rng(SEED)
    xs = points(64, 2, rng)
    mc2 = h.MultiChannel(chans, np.array([0.2, 0.5, 0.3]))
    out["pdf_xs"] = xs
    out["pdf"] = mc2.pdf(xs)
    save("multichannel", **out)


def gen_mixing(h):
    """MixingMarkovUpdate([IS-Metropolis over a MultiChannel, local update]) -- the MC3 sampler
    (base.py:139-179, interfaces/mc3_hmc.py).  Tapes: selector, candidate / momentum / normals, u."""
    out = {}
    D = h.densities
    T = 300
    weights = np.array([0.4, 0.4, 0.2])
    for tag in ("hmc", "rwm", "unif"):
        np.random.seed(SEED + {"hmc": 0, "rwm": 1, "unif": 2}[tag])
        target = D.Camel(2)
        channels = h.MultiChannel([D.Gaussian(2, mu=1 / 3, scale=0.1), D.Gaussian(2, mu=2 / 3, scale=0.1), D.Uniform(2)],
                                  weights.copy())
        is_upd = h.DefaultMetropolis(2, target, channels)
        if tag == "hmc":
            local = h.hamiltonian.HamiltonianUpdate(target, D.Gaussian(2, cov=1.), 10, 0.02)
            local_var = np.array([1., 1.])
        elif tag == "rwm":
            local = h.DefaultMetropolis(2, target, h.proposals.Gaussian(2, cov=np.diag([0.002, 0.003])))
            local_var = np.array([0.002, 0.003])
        else:   # MC3Uniform's local update (mc3.py:100-106); local_var holds delta
            local = h.DefaultMetropolis(2, target, h.proposals.UniformLocal(2, np.array([0.2, 0.3])))
            local_var = np.array([0.2, 0.3])
        mix = h.MixingMarkovUpdate(2, [is_upd, local], [0.4, 0.6])
        x0s = [np.array([0.3, 0.35]), np.array([0.7, 0.6])]
        chains, accs, sels, draws, us = [], [], [], [], []
        for x0 in x0s:
            log = {"sel": [], "cand": None, "p": None, "z": None}
            steps_log = []
            orig_choice = np.random.choice
            np.random.choice = lambda *a, **k: (lambda v: (log["sel"].append(int(v)), v)[1])(orig_choice(*a, **k))
            orig_rvs = channels.rvs
            def rvs(n, return_sizes=False, _o=orig_rvs):
                r = _o(n, return_sizes)
                if n == 1 and not return_sizes:
                    log["cand"] = np.array(r[0], copy=True)
                return r
            channels.rvs = rvs
            if tag == "hmc":
                orig_pp = local.p_dist.proposal
                def pprop(state=None, _o=orig_pp):
                    v = _o(state)
                    log["p"] = np.array(v, copy=True)
                    return v
                local.p_dist.proposal = pprop
            with Tape() as tp:
                orig_next = mix.next_state
                def next_state(state, it, _o=orig_next):
                    nz, nu = len(tp.log["z"]), len(tp.log["rand"])
                    out_state = _o(state, it)
                    sel = log["sel"][-1]
                    nunif = len(tp.log["rand"]) - nu
                    if sel == 0:
                        dr = log["cand"]
                    elif tag == "hmc":
                        dr = log["p"]
                    elif tag == "rwm":
                        dr = np.ravel(tp.log["z"][-1])
                    else:   # the proposal's scalar np.random.rand() comes first, then the accept uniform
                        dr = np.array([float(tp.log["rand"][nu]), np.nan])
                        nunif -= 1
                    assert nunif in (0, 1)
                    steps_log.append((sel, np.array(dr, copy=True), float(tp.log["rand"][-1]) if nunif else np.nan))
                    return out_state
                mix.next_state = next_state
                smp = mix.sample(T, x0, log_every=-1)
                mix.next_state = orig_next
            np.random.choice = orig_choice
            channels.rvs = orig_rvs
            if tag == "hmc":
                local.p_dist.proposal = orig_pp
            chains.append(smp.data)
            accs.append(smp.accepted)
            sels.append([s_[0] for s_ in steps_log])
            draws.append(np.stack([s_[1] for s_ in steps_log]))
            us.append([s_[2] for s_ in steps_log])
        out[tag + "_x0"] = np.stack(x0s)
        out[tag + "_chain"] = np.stack(chains)
        out[tag + "_accepted"] = np.array(accs, dtype=np.int64)
        out[tag + "_sel"] = np.array(sels, dtype=np.int64)
        out[tag + "_draw"] = np.stack(draws)
        out[tag + "_u"] = np.array(us)
        out[tag + "_local_var"] = local_var
    out["weights"] = weights
    save("mixing", **out)


def gen_rambo_diet(h):
    """RamboOnDiet.map / map_inverse (rambo.py:72-151) and MappedDensity (mapping.py)."""
    out = {}
    rng = np.random.default_rng(SEED)
    for n, e_cm in ((3, 50.), (4, 100.), (6, 1000.)):
        m = h.phase_space.RamboOnDiet(e_cm, n)
        xs = rng.random((96, 3 * n - 4))
        p = m.map(xs)
        out["xs_%d" % n] = xs
        out["p_%d" % n] = p
        out["inv_%d" % n] = m.map_inverse(p.copy())
        out["pdf_%d" % n] = m.pdf(xs)
    # MappedDensity: a Gaussian in momentum space seen from the unit cube (2 -> 3 particles, 5 dims)
    m = h.phase_space.RamboOnDiet(50., 3)
    dens = h.densities.Gaussian(12, mu=10., cov=400.)
    md = h.phase_space.MappedDensity(dens, m, norm=2.5)
    xs = points(48, 5, rng)
    out["mapped_xs"] = xs
    out["mapped_pdf"] = md.pdf(xs)
    save("rambo_diet", **out)


def gen_stats(h):
    """stat_tools.effective_sample_size / auto_cov / fd_bins / bin_wise_chi2 (:11-129) on
    reference Metropolis chains over Camel(2) and a 3-D Gaussian.

    Two accommodations for the authoring container, neither touching the arithmetic:
    np.random.uniform is wrapped to log the uniforms behind it (legacy uniform ==
    low + (high - low) * random_sample, asserted), and scipy.stats.chisquare is called with
    sum_check=False -- SciPy >= 1.16 otherwise rejects f_obs/f_exp whose sums differ, which
    the SciPy the reference was written for (>= 1.0) did not."""
    from scipy import stats as sps
    import importlib
    st = importlib.import_module("hepmc.core.util.stat_tools")
    out = {}
    # legacy uniform identity
    np.random.seed(3)
    a = np.random.uniform([0., 1.], [2., 5.], (7, 2))
    np.random.seed(3)
    b = np.array([0., 1.]) + (np.array([2., 5.]) - np.array([0., 1.])) * np.random.random_sample((7, 2))
    assert np.array_equal(a, b), "legacy uniform mismatch"

    orig_uniform, orig_chisq = np.random.uniform, sps.chisquare
    cases = (
        ("camel", h.densities.Camel(2), h.proposals.Gaussian(2, cov=np.diag([0.004, 0.004])), [0.35, 0.3], 6000),
        ("gauss", h.densities.Gaussian(3, mu=0.25, cov=0.7), h.proposals.Gaussian(3, cov=0.5 * np.eye(3)),
         [0.1, -0.2, 0.3], 4000),
    )
    for name, target, prop, x0, T in cases:
        np.random.seed(SEED)
        sample = h.DefaultMetropolis(target.ndim, target, prop).sample(T, x0, log_every=-1)
        sample.target = target
        out[name + "_data"] = np.array(sample.data)
        out[name + "_mean"] = np.asarray(target.mean, dtype=np.float64) * np.ones(target.ndim)
        out[name + "_var"] = np.asarray(target.variance, dtype=np.float64) * np.ones(target.ndim)
        out[name + "_ess"] = st.effective_sample_size(sample, target.mean, target.variance)
        short = sample.data[:300]
        out[name + "_acov300"] = st.auto_cov(short, np.mean(short, 0), np.var(short, 0))
        out[name + "_fd_bins"] = st.fd_bins(sample)
        tape = []

        def uniform(low, high, size):
            u = np.random.random_sample(size)
            tape.append(u.copy())
            return np.asarray(low) + (np.asarray(high) - np.asarray(low)) * u
        np.random.uniform = uniform
        sps.chisquare = lambda f_obs, f_exp: orig_chisq(f_obs, f_exp=f_exp, sum_check=False)
        st.stats.chisquare = sps.chisquare
        try:
            for tag, kw in (("auto", {}), ("fixed", dict(bins=12, min_count=5, int_steps=64))):
                del tape[:]
                res = st.bin_wise_chi2(sample, **kw)
                out["%s_chi2_%s" % (name, tag)] = np.array(res, dtype=np.float64)
                out["%s_chi2_%s_u" % (name, tag)] = np.stack(tape)
        finally:
            np.random.uniform = orig_uniform
            sps.chisquare = orig_chisq
            st.stats.chisquare = orig_chisq
    save("stats", **out)


def gen_stratified(h):
    """StratifiedMC over GridVolumes with uniform and per-bin base counts (stratified.py:68-91,
    stratified_volume.py:127-155); the per-bin np.random.rand draws are taped."""
    out = {}
    cases = (
        ("lin1", 1, 4, 10, {}, 5, lambda x: x),                              # tests/test_integration.py:68-76
        ("camel2", 2, 5, 20, {(1, 1): 100, (3, 3): 60, (0, 4): 8}, 3, None),
    )
    for name, ndim, div, base, counts, multiple, fn in cases:
        np.random.seed(SEED)
        vol = h.GridVolumes(ndim=ndim, divisions=div, default_base_count=base, base_counts=dict(counts))
        target = h.densities.Camel(ndim) if fn is None else None
        with Tape() as t:
            s = h.StratifiedMC(volumes=vol)(target if fn is None else fn, multiple)
        out[name + "_u"] = np.concatenate([np.asarray(r).reshape(-1, ndim) for r in t.log["rand"]])
        out[name + "_sizes"] = np.array(s.sub_sizes, dtype=np.int64)
        out[name + "_integral"] = s.integral
        out[name + "_err"] = s.integral_err
        out[name + "_data"] = s.data
        out[name + "_values"] = s.function_values
        out[name + "_sub_weights"] = np.array(s.sub_weights)
        xs = np.random.default_rng(3).random((64, ndim))
        out[name + "_pdf_xs"] = xs
        out[name + "_pdf"] = vol.pdf(xs)
        out[name + "_total_base_count"] = vol.total_base_count
    save("stratified", **out)


def gen_importance_ideal(h):
    """tests/test_integration.py:44-52 -- the reference's only exact KAT."""
    np.random.seed(SEED)
    dist = h.Distribution.make(lambda x: 2 * x, ndim=1, rvs=lambda size: np.random.rand(size) ** 2)
    s = h.ImportanceMC(dist)(lambda x: x, 1000)
    assert s.integral == 0.5 and s.integral_err == 0.0
    save("importance_ideal", xs=s.data, ys=s.function_values, w=s.weights,
         integral=s.integral, err=s.integral_err)


def main():
    os.makedirs(OUT, exist_ok=True)
    check_mvn()
    h = refshim.load()
    gen_densities(h)
    gen_plain(h)
    gen_vegas(h)
    gen_rambo(h)
    gen_metropolis(h)
    gen_elm_and_hmc(h)
    gen_importance_ideal(h)
    gen_multichannel(h)
    gen_mixing(h)
    gen_rambo_diet(h)
    gen_stats(h)
    gen_stratified(h)
    gen_leapfrog(h)
    gen_ks_reference()


if __name__ == "__main__":
    main()
