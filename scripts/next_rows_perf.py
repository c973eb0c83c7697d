"""Throughput of the SURVEY 8(f) "next" rows that the bench line does not carry: the (MC)^3 mixing kernel
(row 3) and RamboOnDiet map / map_inverse (row 4).  Development aid; numbers quoted in DESIGN.md section 8."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
H.set_outputs("torch")


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


out = {}
D = H.densities
for ndim in (2, 8):
    chans = H.MultiChannel([D.Gaussian(ndim, mu=1 / 3, scale=0.1), D.Gaussian(ndim, mu=2 / 3, scale=0.1), D.Uniform(ndim)])
    tgt = D.Camel(ndim)
    C, T = 200_000, 501
    x0 = torch.full((C, ndim), 1 / 3, device="cuda", dtype=torch.float64)
    # (MC)^3 with a local random-walk update (mc3.BasicMC3) and with a local HMC update (mc3.MC3Hamilton, exact gradient)
    rw = H.mc3.BasicMC3(tgt, chans, H.DefaultMetropolis(ndim, tgt, H.proposals.Gaussian(ndim, cov=0.002 * np.eye(ndim))), beta=0.5)
    t = timed(lambda: rw.mixing_sampler.sample(T, x0, store_chain=False))
    s = rw.mixing_sampler.sample(T, x0, store_chain=False)
    out["mc3_rwm_%dd" % ndim] = {"chain_steps_per_s": C * (T - 1) / t, "acceptance": float(s.accepted.double().mean()) / T}
    hm = H.mc3.MC3Hamilton(tgt, chans, np.ones(ndim), 10, 0.02, beta=0.5)
    Th = 101
    t = timed(lambda: hm.mixing_sampler.sample(Th, x0, store_chain=False))
    s = hm.mixing_sampler.sample(Th, x0, store_chain=False)
    out["mc3_hmc10_%dd" % ndim] = {"chain_steps_per_s": C * (Th - 1) / t, "acceptance": float(s.accepted.double().mean()) / Th}
for n in (4, 6):
    m = H.phase_space.RamboOnDiet(100., n)
    N = 1 << 24
    xs = torch.rand((N, 3 * n - 4), device="cuda", dtype=torch.float64)
    t = timed(lambda: m.map(xs))
    p = m.map(xs)
    t2 = timed(lambda: m.map_inverse(p))
    out["rambo_diet_n%d" % n] = {"map_points_per_s": N / t, "map_GBs": N * 8 * (3 * n - 4 + 4 * n) / t / 1e9,
                                 "map_inverse_points_per_s": N / t2}
print(json.dumps(out, indent=1))
