"""One launch of each non-headline kernel (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
H.set_outputs("torch")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "rambo"):
    m = H.phase_space.Rambo(100., 4)
    buf = torch.empty((1 << 25, 16), dtype=torch.float64, device="cuda")
    for _ in range(2):
        m.generate(1 << 25, out=buf)
if which in ("all", "rwm"):
    C = 1 << 20
    met = H.DefaultMetropolis(2, H.densities.Banana(2), H.proposals.Gaussian(2, cov=10 * np.eye(2)))
    x0 = torch.tensor([[0., 10.]], device="cuda", dtype=torch.float64).repeat(C, 1)
    for _ in range(2):
        met.sample(201, x0, store_chain=False)
if which in ("all", "hmc"):
    rs = np.random.default_rng(0)
    nodes = 1024
    basis = H.surrogate.GaussianBasis(8)
    params = (rs.random((nodes, 8)), 0.01 + 0.99 * rs.random(nodes), None)
    w = 0.01 * rs.standard_normal(nodes)
    for prec in ("f64", "f32"):
        tgt = H.densities.Camel(8)
        H.surrogate.attach_surrogate_gradient(tgt, basis, params, 0.0, w)
        upd = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8), 20, 0.01, precision=prec)
        x0h = torch.full((100000, 8), 1 / 3, device="cuda", dtype=torch.float64)
        for _ in range(2):
            upd.sample(2, x0h, store_chain=False)
torch.cuda.synchronize()
print("done")
