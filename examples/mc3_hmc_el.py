"""(MC)^3 with an extreme-learning-machine surrogate gradient: the reference's
interfaces/mc3_hmc_el.py recipe on the 4-D camel, one chain and an ensemble."""
import numpy as np
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from anywhere
import hepmc_b200 as hepmc
from hepmc_b200 import densities, interfaces

hepmc.seed(1234)
ndim = 4
target = densities.Camel(ndim)
sampler, initial = interfaces.mc3_hmc_el.get_sampler(
    target, ndim, 0.4, centers=[1 / 3, 2 / 3], widths=[0.1, 0.1], beta=0.3,
    nintegral=20_000, mass=1., steps=10, step_size=0.02, el_size=200)

one = sampler.sample(5_000, initial)
print(one)                                  # mean, variance, bin-wise chi^2, effective sample size, acceptance

many = sampler.sample(500, np.tile(initial, (20_000, 1)), store_chain=False)
pts = many.data[:, 0]
print("ensemble of 20000 chains: mean %s  variance %s" % (np.round(pts.mean(axis=0), 3), np.round(pts.var(axis=0), 4)))
print("target                  : mean %s  variance %.4f" % (np.full(ndim, 0.5), 0.005 + 1 / 36))
