"""Pin the NumPy oracle against outputs of the RUNNING REFERENCE (tests/golden,
made by oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, assert_close
from oracle import hepmc_oracle as O

G = load_golden


@pytest.mark.parametrize("ndim", [1, 2, 3, 8, 16])
def test_densities(ndim):
    g = G("densities")
    xs = g["xs_%d" % ndim]
    for tag, bounded in (("camel", True), ("ucamel", False)):
        assert_close(O.camel_pdf(xs, bounded=bounded), g["%s_pdf_%d" % (tag, ndim)], 1e-13, msg=tag + " pdf")
        assert_close(O.camel_pdf_gradient(xs, bounded=bounded), g["%s_pdfgrad_%d" % (tag, ndim)], 1e-13, msg=tag + " pdfgrad")
        assert_close(O.camel_pot(xs, bounded=bounded), g["%s_pot_%d" % (tag, ndim)], 1e-13, msg=tag + " pot")
        with np.errstate(all="ignore"):
            assert_close(O.camel_pot_gradient(xs, bounded=bounded), g["%s_potgrad_%d" % (tag, ndim)], 1e-12, atol=1e-12, msg=tag + " potgrad")
    assert_close(O.gaussian_pdf(xs, 0.25, 0.7), g["gauss_pdf_%d" % ndim], 1e-13)
    assert_close(O.gaussian_pot(xs, 0.25, 0.7), g["gauss_pot_%d" % ndim], 1e-13)
    assert_close(O.gaussian_pot_gradient(xs, 0.25, 0.7), g["gauss_potgrad_%d" % ndim], 1e-13, atol=1e-15)
    mu, cov = g["gaussd_mu_%d" % ndim], g["gaussd_cov_%d" % ndim]
    assert_close(O.gaussian_pdf(xs, mu, cov), g["gaussd_pdf_%d" % ndim], 1e-13)
    assert_close(O.gaussian_pot(xs, mu, cov), g["gaussd_pot_%d" % ndim], 1e-13)
    assert_close(O.gaussian_pot_gradient(xs, mu, cov), g["gaussd_potgrad_%d" % ndim], 1e-13, atol=1e-15)
    assert_close(O.uniform_pdf(xs), g["uniform_pdf_%d" % ndim], 0)
    assert_close(O.uniform_pot(xs), g["uniform_pot_%d" % ndim], 0)


def test_density_tail_and_banana():
    g = G("densities")
    assert_close(O.camel_pdf(g["xs_tail"], bounded=False), g["ucamel_pdf_tail"], 2e-13)
    for ndim in (2, 3):
        xs = g["banana_xs_%d" % ndim]
        assert_close(O.banana_pdf(xs), g["banana_pdf_%d" % ndim], 1e-13)
        assert_close(O.banana_pot(xs), g["banana_pot_%d" % ndim], 1e-13)
    assert_close(O.banana_pot_gradient(g["banana_xs_2"]), g["banana_potgrad_2"], 1e-13)
    assert_close(O.camel_pdf(g["xs_3"]), g["camel_call_3"], 1e-13)


@pytest.mark.parametrize("ndim", [2, 16])
def test_plain_mc(ndim):
    g = G("plain_mc")
    integral, err, f = O.plain_mc(O.camel_pdf, g["data_%d" % ndim])
    assert_close(f, g["f_%d" % ndim], 1e-13)
    assert_close(integral, g["integral_%d" % ndim], 1e-13)
    assert_close(err, g["err_%d" % ndim], 1e-12)


@pytest.mark.parametrize("tag", ["2d", "16d", "3d_vw"])
def test_vegas(tag):
    g = G("vegas")
    ndim, K, c, n, iters, vw = g["params_" + tag]
    ndim, K, n, iters = int(ndim), int(K), int(n), int(iters)
    r = O.vegas(O.camel_pdf, ndim, K, g["idx_" + tag], g["u_" + tag], c=c,
                var_weighted=bool(vw), chi=True)
    assert_close(r["bounds_history"], g["bounds_" + tag], 1e-12, atol=1e-15, msg="bounds")
    assert_close(r["integral"], g["integral_" + tag], 1e-12)
    assert_close(r["integral_err"], g["err_" + tag], 1e-11)
    assert_close(r["chi2"], g["chi2_" + tag], 1e-10)
    assert_close(r["data"], g["data_" + tag], 1e-13)
    assert_close(r["function_values"], g["f_" + tag], 1e-12)
    assert_close(r["weights"], g["w_" + tag], 1e-12)


def test_grid_pdf_and_importance():
    g = G("vegas")
    b = g["gridpdf_bounds"]
    assert_close(O.grid_pdf(b, g["gridpdf_xs"]), g["gridpdf_pdf"], 1e-13)
    x = O.grid_sample(b, g["imp_idx"], g["imp_u"]).T
    assert_close(x, g["imp_data"], 1e-14)
    w = O.grid_pdf(b, x)
    assert_close(w, g["imp_w"], 1e-12)
    f = O.camel_pdf(x)
    integral, err = O.importance_mc(f, w)
    assert_close(integral, g["imp_integral"], 1e-12)
    assert_close(err, g["imp_err"], 1e-11)


def test_importance_ideal_kat():
    """reference tests/test_integration.py:44-52: exact 0.5 +- 0.0."""
    g = G("importance_ideal")
    integral, err = O.importance_mc(g["ys"][:, 0] if g["ys"].ndim == 2 else g["ys"], g["w"])
    assert integral == 0.5 and err == 0.0


@pytest.mark.parametrize("tag,e_cm,npart", [("4_100", 100., 4), ("3_50", 50., 3),
                                            ("6_1000", 1000., 6), ("2_10", 10., 2)])
def test_rambo(tag, e_cm, npart):
    g = G("rambo")
    p = O.rambo_map(g["xs_" + tag], e_cm, npart)
    assert_close(p, g["p_" + tag], 1e-12, atol=1e-12 * e_cm)
    assert_close(O.rambo_pdf(e_cm, npart), g["pdf_" + tag], 1e-15)
    # physics invariants: sum p = (E,0,0,0), p^2 = 0
    P = p.reshape(-1, npart, 4)
    tot = P.sum(axis=1)
    assert np.allclose(tot[:, 0], e_cm, rtol=1e-13)
    assert np.all(np.abs(tot[:, 1:]) < 1e-12 * e_cm)
    m2 = P[:, :, 0] ** 2 - np.sum(P[:, :, 1:] ** 2, axis=2)
    assert np.all(np.abs(m2) < 1e-9 * e_cm ** 2)


def test_rambo_pdf_kat():
    assert O.rambo_pdf(100., 4) == pytest.approx(3.0961473055871514e-08, rel=1e-15)


def test_rambo_importance():
    g = G("rambo")
    p = O.rambo_map(g["imp_xs"], 100., 4)
    assert_close(p, g["imp_data"], 1e-12, atol=1e-10)
    f = p[:, 0] * p[:, 4] / 1e3 + p[:, 3] ** 2 / 1e3
    assert_close(f, g["imp_f"], 1e-11)
    integral, err = O.importance_mc(f, float(g["imp_w"]))
    assert_close(integral, g["imp_integral"], 1e-11)
    assert_close(err, g["imp_err"], 1e-10)


def _fill(u):
    """NaN entries = uniforms the reference never drew (accept >= 1)."""
    return np.where(np.isnan(u), 2.0, u)


@pytest.mark.parametrize("tag", ["banana", "camel", "gauss"])
def test_metropolis(tag):
    g = G("metropolis")
    pdf = {"banana": O.banana_pdf, "camel": O.camel_pdf,
           "gauss": lambda x: O.gaussian_pdf(x, 0.25, 0.7)}[tag]
    chain, acc = O.metropolis_chains(pdf, g[tag + "_x0"], g[tag + "_z"], _fill(g[tag + "_u"]),
                                     np.sqrt(g[tag + "_cov"]))
    assert np.array_equal(chain, g[tag + "_chain"])
    assert np.array_equal(acc, g[tag + "_accepted"])


def test_elm():
    g = G("elm")
    for tag in ("g64", "g100"):
        c, w, q = g[tag + "_centers"], g[tag + "_widths"], g[tag + "_q"]
        assert_close(O.elm_gauss_output_matrix(q, c, w), g[tag + "_outmat"], 1e-13)
        assert_close(O.elm_gauss_eval(q, c, w, g[tag + "_bias"], g[tag + "_weights"]), g[tag + "_eval"], 1e-9, atol=1e-9)
        assert_close(O.elm_gauss_eval_gradient(q, c, w, g[tag + "_weights"]), g[tag + "_grad"], 1e-9, atol=1e-7)
    assert_close(O.elm_trig_eval(g["g64_q"], g["t48_biases"], g["t48_inweights"], g["t48_bias"], g["t48_weights"]),
                 g["t48_eval"], 1e-9, atol=1e-6)
    assert_close(O.elm_trig_eval_gradient(g["g64_q"], g["t48_biases"], g["t48_inweights"], g["t48_weights"]),
                 g["t48_grad"], 1e-9, atol=1e-6)


@pytest.mark.parametrize("tag", ["exact", "elm", "banana"])
def test_hmc(tag):
    g = G("hmc")
    e = G("elm")
    steps, eps, mass = g[tag + "_params"]
    if tag == "exact":
        pdf, grad = O.camel_pdf, O.camel_pot_gradient
    elif tag == "elm":
        pdf = O.camel_pdf
        grad = lambda q: O.elm_gauss_eval_gradient(q, e["g100_centers"], e["g100_widths"], e["g100_weights"])
    else:
        pdf, grad = O.banana_pdf, O.banana_pot_gradient
    chain, acc = O.hmc_chains(pdf, grad, g[tag + "_x0"], g[tag + "_p"], _fill(g[tag + "_u"]),
                              mass, eps, int(steps))
    assert np.array_equal(acc, g[tag + "_accepted"])
    assert_close(chain, g[tag + "_chain"], 1e-9, atol=1e-9)


@pytest.mark.parametrize("tag", ["exact", "elm", "banana", "banana0"])
def test_leapfrog(tag):
    """O.leapfrog against the reference's HamiltonLeapfrog.__call__ (simulation.py:26-46), bit for bit."""
    g = G("leapfrog")
    e = G("elm")
    steps, eps = g[tag + "_params"]
    base = tag.rstrip("0")
    if base == "exact":
        grad = O.camel_pot_gradient
    elif base == "elm":
        grad = lambda q: O.elm_gauss_eval_gradient(q, e["g100_centers"], e["g100_widths"], e["g100_weights"])
    else:
        grad = O.banana_pot_gradient
    kin = lambda p: O.gaussian_pot_gradient(p, 0.0, g[base + "_mass"])
    q, p = O.leapfrog(grad, kin, g[tag + "_q0"], g[tag + "_p0"], eps, int(steps))
    assert np.array_equal(q, g[tag + "_q"]) and np.array_equal(p, g[tag + "_p"])


@pytest.mark.parametrize("tag", ["exact", "elm", "banana"])
def test_hmc_single_transitions(tag):
    """Every transition of the reference chains on its own (T = 2): the oracle is bit-exact per transition."""
    g = G("hmc")
    e = G("elm")
    steps, eps, mass = g[tag + "_params"]
    if tag == "exact":
        pdf, grad = O.camel_pdf, O.camel_pot_gradient
    elif tag == "elm":
        pdf = O.camel_pdf
        grad = lambda q: O.elm_gauss_eval_gradient(q, e["g100_centers"], e["g100_widths"], e["g100_weights"])
    else:
        pdf, grad = O.banana_pdf, O.banana_pot_gradient
    chain = g[tag + "_chain"]
    nd = chain.shape[2]
    starts, expect = chain[:, :-1].reshape(-1, nd), chain[:, 1:].reshape(-1, nd)
    c2, acc = O.hmc_chains(pdf, grad, starts, g[tag + "_p"].reshape(-1, 1, nd), _fill(g[tag + "_u"]).reshape(-1, 1),
                           mass, eps, int(steps))
    assert np.array_equal(c2[:, 1], expect)


def _mc_channels(g):
    return [("gauss", 1 / 3, 0.01), ("gauss", g["chan_mu1"], g["chan_cov1"]), ("uniform", 0., 1.)]


def test_multichannel():
    g = G("multichannel")
    chans = _mc_channels(g)
    n1, n2, n3 = [int(v) for v in g["n_iter"]]
    tapes = [(g["xs_%d" % j], g["sizes_%d" % j]) for j in range(n1 + n2 + n3)]
    r = O.multichannel_mc(O.camel_pdf, chans, np.ones(3) / 3, tapes, n1, n2, n3)
    assert_close(r["integral"], g["integral"], 1e-13)
    assert_close(r["integral_err"], g["err"], 1e-12)
    assert_close(r["alpha"], g["alpha_final"], 1e-13)
    assert_close(O.multichannel_pdf(chans, [0.2, 0.5, 0.3], g["pdf_xs"]), g["pdf"], 1e-14)


@pytest.mark.parametrize("tag,kind", [("hmc", 1), ("rwm", 0), ("unif", 2)])
def test_mixing_chains(tag, kind):
    g = G("mixing")
    chans = [("gauss", 1 / 3, 0.01), ("gauss", 2 / 3, 0.01), ("uniform", 0., 1.)]
    chain, acc = O.mixing_chains(O.camel_pdf, chans, g["weights"], kind, g[tag + "_local_var"], g[tag + "_x0"],
                                 g[tag + "_sel"], g[tag + "_draw"], _fill(g[tag + "_u"]),
                                 pot_grad_fn=O.camel_pot_gradient, step_size=0.02, steps=10)
    assert np.array_equal(acc, g[tag + "_accepted"])
    assert_close(chain, g[tag + "_chain"], 1e-10, atol=1e-12)


def test_uniform_local_chains():
    """DefaultMetropolis + proposals.UniformLocal (uniform_local.py:27-29), one start in a corner."""
    g = G("metropolis")
    chain, acc = O.uniform_local_chains(O.camel_pdf, g["ulocal_x0"], g["ulocal_v"], _fill(g["ulocal_u"]), g["ulocal_delta"])
    assert np.array_equal(acc, g["ulocal_accepted"])
    assert np.array_equal(chain, g["ulocal_chain"])


@pytest.mark.parametrize("n,e_cm", [(3, 50.), (4, 100.), (6, 1000.)])
def test_rambo_diet(n, e_cm):
    g = G("rambo_diet")
    p = O.rambo_diet_map(g["xs_%d" % n], e_cm, n)
    # the reference's root is brentq with xtol = 2e-12: parity to ~1e-11 of e_cm
    assert_close(p, g["p_%d" % n], 1e-10, atol=1e-11 * e_cm)
    assert_close(O.rambo_diet_inverse(g["p_%d" % n], e_cm, n), g["inv_%d" % n], 1e-12, atol=1e-13)
    assert_close(O.rambo_pdf(e_cm, n), g["pdf_%d" % n], 1e-15)


def test_philox_kat():
    """Random123 known-answer vectors for Philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        out = O.philox4x32(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
        assert tuple(int(o[0]) for o in out) == exp


def test_philox7_kat():
    """Random123 known-answer vectors for Philox4x32-7 (kat_vectors: philox4x32 7 ...)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a))]
    for ctr, key, exp in kat:
        out = O.philox4x32(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1], rounds=7)
        assert tuple(int(o[0]) for o in out) == exp


def test_32bit_draw_conventions():
    """u32 / vegas_bin_frac32: exact values, the kernels' exponent trick, the joint (bin, offset) law."""
    w = np.array([0, 1, 0x80000000, 0xffffffff, 0x12345678], dtype=np.uint32)
    u = O.u32(w)
    assert u[0] == 2.0 ** -33 and u[3] == 1.0 - 2.0 ** -33 and np.all((u > 0) & (u < 1))
    # as_double(0x41300000 : w) - (2^20 - 2^-33), csrc/integrate_kernels.cuh u32()
    trick = ((np.uint64(0x41300000) << np.uint64(32)) | w.astype(np.uint64)).view(np.float64) - (2.0 ** 20 - 2.0 ** -33)
    assert np.array_equal(trick, u)
    for k in (1, 2, 15, 50, 255):
        rs = np.random.RandomState(k)
        ww = np.concatenate([w, rs.randint(0, 2 ** 32, 20000, dtype=np.uint64).astype(np.uint32)])
        b, f = O.vegas_bin_frac32(ww, k)
        assert b.min() >= 0 and b.max() < k and np.all((f > 0) & (f < 1))
        # (bin + offset) / K reproduces the uniform the word stands for, shifted by half a quantum of the offset
        t = (b + f) / k
        assert np.allclose(t, (ww.astype(np.float64) + 0.5 / k) * 2.0 ** -32, rtol=0, atol=2e-16)
    # all three modes of the draw helpers agree on shapes and differ in values
    idx = np.arange(50)
    d52 = O.philox_vegas_draws(7, 0, idx, 5, 9, "u52r10")
    d32 = O.philox_vegas_draws(7, 0, idx, 5, 9, "u32r10")
    d7 = O.philox_vegas_draws(7, 0, idx, 5, 9, "u32r7")
    assert d52[0].shape == d32[0].shape == d7[0].shape == (5, 50)
    assert not np.array_equal(d52[1], d32[1]) and not np.array_equal(d32[1], d7[1])
    u = O.philox_uniforms(7, 0, idx, 6, "u32r10")
    w4 = O.philox_words(7, 0, idx, 2)
    assert np.array_equal(u[:, 5], O.u32(w4[:, 1, 1]))


def test_uniform_conventions():
    a = np.array([0, 0xffffffff, 0x12345678], dtype=np.uint32)
    b = np.array([0, 0xffffffff, 0x9abcdef0], dtype=np.uint32)
    u = O.u52(a, b)
    assert u[0] == 2.0 ** -53 and u[1] == 1.0 - 2.0 ** -53
    assert np.all((u > 0) & (u < 1))
    # bit-trick form used by the kernels
    k = ((a.astype(np.uint64) >> np.uint64(12)) << np.uint64(32)) | b.astype(np.uint64)
    trick = (k | np.uint64(0x3FF0000000000000)).view(np.float64) - (1.0 - 2.0 ** -53)
    assert np.array_equal(trick, u)
    wa = np.array([0, 0xffffffff, 0x80000000, 0x12345678], dtype=np.uint32)
    wb = np.array([0, 0xffffffff, 0, 0x9abcdef0], dtype=np.uint32)
    for K in (1, 2, 15, 16, 50, 255):
        bins, frac = O.vegas_bin_frac(wa, wb, K)
        assert bins.max() <= K - 1 and bins.min() >= 0 and np.all((frac > 0) & (frac < 1))
        t = K * ((wa.astype(object) << 32) + wb.astype(object))      # exact big ints
        assert [int(v) >> 64 for v in t] == list(bins)


def test_ks_restatement():
    g = G("camel2d_reference_hist")
    from oracle.stats_oracle import ks
    for fam in ("cv", "hmc"):
        assert_close(np.array(ks(g["chain_hist_" + fam], g["truth"])), g["chain_ks_" + fam], 1e-13)


def _stats_cases():
    return (("camel", lambda xs: O.camel_pdf(xs)),
            ("gauss", lambda xs: O.gaussian_pdf(xs, np.full(3, 0.25), np.full(3, 0.7))))


def test_effective_sample_size_restatement():
    """stat_tools.py:17-70 on reference Metropolis chains."""
    g = G("stats")
    for name, _ in _stats_cases():
        data = g[name + "_data"]
        ess, stop = O.effective_sample_size(data, g[name + "_mean"], g[name + "_var"])
        assert_close(ess, g[name + "_ess"], 1e-12)
        assert np.all(stop >= 1)
        short = data[:300]
        assert_close(O.auto_cov(short, np.mean(short, 0), np.var(short, 0)), g[name + "_acov300"], 1e-13, atol=1e-18)


def test_bin_wise_chi2_restatement():
    """stat_tools.py:73-129 with the reference's uniforms replayed."""
    g = G("stats")
    for name, pdf in _stats_cases():
        data = g[name + "_data"]
        assert np.array_equal(O.fd_bins(data), g[name + "_fd_bins"])
        for tag, kw in (("auto", {}), ("fixed", dict(bins=12, min_count=5, int_steps=64))):
            red, p, n, _ = O.bin_wise_chi2(data, pdf, u_tape=g["%s_chi2_%s_u" % (name, tag)], **kw)
            ref = g["%s_chi2_%s" % (name, tag)]
            assert n == int(ref[2])
            assert_close(red, ref[0], 1e-12)
            assert_close(p, ref[1], 1e-9)


def test_spare_bit_uniform():
    a = np.array([0, 0xffffffff, 0x12345abc], dtype=np.uint32)
    c = np.array([0, 0xffffffff, 0x9abcd123], dtype=np.uint32)
    u = O.u24_spare(a, c)
    assert u[0] == 2.0 ** -25 and u[1] == 1.0 - 2.0 ** -25 and u[2] == (0xabc * 4096 + 0x123 + 0.5) * 2.0 ** -24


@pytest.mark.parametrize("name,ndim,div,base,counts,multiple", [
    ("lin1", 1, 4, 10, {}, 5), ("camel2", 2, 5, 20, {(1, 1): 100, (3, 3): 60, (0, 4): 8}, 3)])
def test_stratified_restatement(name, ndim, div, base, counts, multiple):
    """stratified.py:68-91 + stratified_volume.py:67-83,127-155 with the reference's draws replayed."""
    g = G("stratified")
    bounds = [np.linspace(0, 1, div + 1) for _ in range(ndim)]
    fn = (lambda xs: xs[:, 0]) if name == "lin1" else O.camel_pdf
    sizes = g[name + "_sizes"]
    tape = np.split(g[name + "_u"], np.cumsum(sizes)[:-1])
    r = O.stratified_mc(fn, bounds, tape, multiple, base, counts)
    assert np.array_equal(r["sizes"], sizes)
    assert_close(r["integral"], g[name + "_integral"], 1e-13)
    assert_close(r["integral_err"], g[name + "_err"], 1e-12)
    assert_close(r["data"], g[name + "_data"], 1e-15)
    assert_close(O.grid_pdf_with_counts(bounds, g[name + "_pdf_xs"], base, counts), g[name + "_pdf"], 1e-14)
