"""Measured: error of ONE HMC transition against the reference (per tag), and how the deviation of a
replayed chain grows with the step index (trajectory amplification).  Fixtures: tests/golden/hmc.npz."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hepmc_b200 as H

g = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "hmc.npz")))
e = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "elm.npz")))
lf = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "leapfrog.npz")))
fill = lambda u: np.where(np.isnan(u), 2.0, u)
out = {}
for tag in ("exact", "elm", "banana"):
    steps, eps, mass = g[tag + "_params"]
    chain = g[tag + "_chain"]
    S, T, nd = chain.shape
    target = H.densities.Banana(2) if tag == "banana" else H.densities.Camel(nd)
    if tag == "elm":
        H.surrogate.attach_surrogate_gradient(target, H.surrogate.GaussianBasis(nd), (e["g100_centers"], e["g100_widths"], None),
                                              float(e["g100_bias"]), e["g100_weights"])
    upd = H.hamiltonian.HamiltonianUpdate(target, H.densities.Gaussian(nd, cov=mass), int(steps), float(eps))
    starts, expect = chain[:, :-1].reshape(-1, nd), chain[:, 1:].reshape(-1, nd)
    s = upd.sample(2, starts, replay=(g[tag + "_p"].reshape(-1, 1, nd), fill(g[tag + "_u"]).reshape(-1, 1)))
    rel = np.abs(s.data[:, 1] - expect) / np.maximum(np.abs(expect), 1e-300)
    out[tag + "_single_transition_max_rel"] = float(rel.max())
    full = upd.sample(T, g[tag + "_x0"], replay=(g[tag + "_p"], fill(g[tag + "_u"])))
    dev = np.abs(full.data - chain).max(axis=(0, 2))
    out[tag + "_chain_dev_at_t"] = {int(t): float(dev[t]) for t in (1, 5, 10, 20, 40, T - 1)}
    sim = H.hamiltonian.HamiltonLeapfrog(target.pot_gradient, H.densities.Gaussian(nd, cov=float(lf[tag + "_mass"][0])).pot_gradient,
                                         float(lf[tag + "_params"][1]), int(lf[tag + "_params"][0]))
    q, p = sim(lf[tag + "_q0"], lf[tag + "_p0"])
    out[tag + "_leapfrog_max_rel_q"] = float(np.max(np.abs(q - lf[tag + "_q"]) / np.abs(lf[tag + "_q"])))
    out[tag + "_leapfrog_max_abs_p"] = float(np.max(np.abs(p - lf[tag + "_p"])))
    if tag == "elm":
        grad = target.pot_gradient
        out["elm_gradient_cancellation"] = grad.cancellation(lf["elm_q0"])
print(json.dumps(out, indent=1))
