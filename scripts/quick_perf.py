"""Quick device-side timings of every hot-path kernel (development aid, not the bench)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hepmc_b200 as H
from hepmc_b200 import _lib

H.set_outputs("torch")


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


out = {}
out["fp64_probe_tflops"] = _lib.fp64_probe(1 << 15) / 1e12
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 8
mc = H.VegasMC(16, divisions=15)
t = timed(lambda: mc(H.densities.Camel(16), n, 2, keep_samples=False))
out["vegas16_pts_per_s"] = 2 * n / t
mc2 = H.VegasMC(2, divisions=50)
t = timed(lambda: mc2(H.densities.Camel(2), 10 ** 6, 10, keep_samples=False))
out["vegas2_1e6x10_pts_per_s"] = 1e7 / t
t = timed(lambda: H.PlainMC(16)(H.densities.Camel(16), n, keep_samples=False))
out["plain16_pts_per_s"] = n / t
m = H.phase_space.Rambo(100., 4)
nr = 1 << 25
buf = torch.empty((nr, 16), dtype=torch.float64, device="cuda")
t = timed(lambda: m.generate(nr, out=buf))
out["rambo_pts_per_s"] = nr / t
out["rambo_write_GBs"] = nr * 128 / t / 1e9
C, T = 1 << 20, 1001
met = H.DefaultMetropolis(2, H.densities.Banana(2), H.proposals.Gaussian(2, cov=10 * np.eye(2)))
x0 = torch.tensor([[0., 10.]], device="cuda", dtype=torch.float64).repeat(C, 1)
t = timed(lambda: met.sample(T, x0, store_chain=False))
out["rwm_steps_per_s"] = C * (T - 1) / t
rs = np.random.default_rng(0)
for nodes in (100, 1024):
    basis = H.surrogate.GaussianBasis(8)
    params = (rs.random((nodes, 8)), 0.01 + 0.99 * rs.random(nodes), None)
    w = rs.standard_normal(nodes)
    for prec in ("f64", "f32"):
        tgt = H.densities.Camel(8)
        H.surrogate.attach_surrogate_gradient(tgt, basis, params, 0.0, w)
        upd = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8), 20, 0.01, precision=prec)
        Ch, Th = 100000, 3
        x0h = torch.full((Ch, 8), 1 / 3, device="cuda", dtype=torch.float64)
        t = timed(lambda: upd.sample(Th, x0h, store_chain=False), reps=2)
        out["hmc_elm%d_%s_chainsteps_per_s" % (nodes, prec)] = Ch * (Th - 1) / t
print(json.dumps(out, indent=1))
