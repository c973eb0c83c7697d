import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
from hepmc_b200 import _lib
H.set_outputs("torch")
lib = _lib.load()
rs = np.random.default_rng(0)
x0h = torch.full((148 * 256, 8), 1 / 3, device="cuda", dtype=torch.float64)
PREC = sys.argv[1] if len(sys.argv) > 1 else "tc"
for mode in (0,):
    for nodes in (512, 1024):
        basis = H.surrogate.GaussianBasis(8)
        params = (rs.random((nodes, 8)), 0.01 + 0.99 * rs.random(nodes), None)
        w = 0.01 * rs.standard_normal(nodes)
        tgt = H.densities.Camel(8)
        H.surrogate.attach_surrogate_gradient(tgt, basis, params, 0.0, w)
        upd = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8), 20, 0.01, precision=PREC)
        assert lib.hepmc_tc_debug(mode) == 0, "build the library with make EXTRA=-DHEPMC_TC_DEBUG"
        upd.sample(2, x0h, store_chain=False); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); upd.sample(11, x0h, store_chain=False); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print(PREC, "mode %2d nodes %4d: %.1f us per chain-step -> %.0f cycles per gradient eval" % (mode, nodes, ms * 1e3, ms * 1e-3 * 1.965e9 / 21 / 2))
