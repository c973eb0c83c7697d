import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
H.set_outputs("torch")
rs = np.random.default_rng(0)
nodes = 1024
basis = H.surrogate.GaussianBasis(8)
params = (rs.random((nodes, 8)), 0.01 + 0.99 * rs.random(nodes), None)
w = 0.01 * rs.standard_normal(nodes)
tgt = H.densities.Camel(8)
H.surrogate.attach_surrogate_gradient(tgt, basis, params, 0.0, w)
upd = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8), 20, 0.01, precision="tc")
x0h = torch.full((148 * 256, 8), 1 / 3, device="cuda", dtype=torch.float64)
for _ in range(2):
    upd.sample(2, x0h, store_chain=False)
torch.cuda.synchronize()
print("done")
