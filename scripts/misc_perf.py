"""Callers either side of the kernels at the reference's default settings (development script)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
H.set_outputs("torch")

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts)

n = 10_000_000
rambo = H.densities.Rambo(4, 100.)
fn = lambda *ps: torch.ones_like(ps[0])      # examples/rambo.py integrates a constant-like matrix element
t = timed(lambda: H.ImportanceMC(rambo)(fn, n))
print("ImportanceMC(densities.Rambo(4, 100))(fn, 1e7)          : %.3e points/s" % (n / t))
g = H.densities.Gaussian(8, mu=0.5, cov=0.01)
t = timed(lambda: g.rvs(n))
print("densities.Gaussian(8).rvs(1e7)                           : %.3e points/s" % (n / t))
t = timed(lambda: H.ImportanceMC(g)(H.densities.Camel(8), n))
print("ImportanceMC(Gaussian(8))(Camel(8), 1e7)                 : %.3e points/s" % (n / t))
camel = H.densities.Camel(2)
bound = float(camel.pdf(np.array([[1 / 3, 1 / 3]]))[0]) * 1.01
t = timed(lambda: H.AcceptRejectSampler(camel, bound, ndim=2).sample(2_000_000))
print("AcceptRejectSampler(Camel(2)).sample(2e6)                : %.3e accepted points/s" % (2e6 / t))
xs = torch.rand(n, 8, device="cuda", dtype=torch.float64)
c8 = H.densities.Camel(8)
for name, f in (("pdf", c8.pdf), ("pot", c8.pot), ("pot_gradient", c8.pot_gradient)):
    t = timed(lambda: f(xs))
    print("Camel(8).%-12s on 1e7 points                      : %.3e points/s" % (name, n / t))
