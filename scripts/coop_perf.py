"""Cooperative (CTA per chain) vs thread-per-chain HMC kernel at several ensemble sizes (development
script; run with HEPMC_HMC_COOP_MAX=0 for the thread-per-chain kernel)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
H.set_outputs("torch")
rs = np.random.default_rng(0)
for nodes in (100, 1024):
    basis = H.surrogate.GaussianBasis(8)
    params = (rs.random((nodes, 8)), 0.01 + 0.99 * rs.random(nodes), None)
    w = 0.01 * rs.standard_normal(nodes)
    for prec in ("f64", "f32"):
        for C in (1, 64, 1024, 4096, 16384, 65536):
            tgt = H.densities.Camel(8)
            H.surrogate.attach_surrogate_gradient(tgt, basis, params, 0.0, w)
            upd = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8), 20, 0.01, precision=prec)
            x0 = torch.full((C, 8), 1 / 3, device="cuda", dtype=torch.float64)
            T = max(3, min(201, int(2e6 / (C * nodes / 100))))
            upd.sample(3, x0, store_chain=False); torch.cuda.synchronize()
            t0 = time.perf_counter(); upd.sample(T, x0, store_chain=False); torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print("nodes %4d %s C %6d: %.3e chain-steps/s total, %.3e per chain" % (nodes, prec, C, C * (T - 1) / dt, (T - 1) / dt))
