"""Accuracy of the in-kernel exp / log / sincos against NumPy, through the kernels that use them."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
rs = np.random.default_rng(0)
# exp through Gaussian(1).pdf: pdf = exp(-0.5*(log2pi + x^2))
x = np.concatenate([rs.uniform(-38, 38, 200000), rs.uniform(-1, 1, 100000)])[:, None]
got = H.densities.Gaussian(1).pdf(x)
ref = np.exp(-0.5 * (np.log(2 * np.pi) + (x[:, 0] * 1.0) ** 2))
rel = np.abs(got - ref) / ref
print("exp (via Gaussian pdf): max rel err %.3e, denormal region ok: %s" % (rel[ref > 1e-300].max(), np.all(np.isfinite(got))))
# log + sincos through RAMBO n=2 map on crafted inputs
u = rs.random((300000, 8))
p = H.phase_space.Rambo(100., 2).map(u)
from oracle import hepmc_oracle as O
pr = O.rambo_map(u, 100., 2)
print("rambo map max abs err / E: %.3e" % (np.abs(p - pr).max() / 100.))
