"""Random-walk Metropolis on the banana density: one chain like the reference's
examples/rwm_banana.py, then an ensemble of 100 000 chains in one kernel."""
import numpy as np
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from anywhere
import hepmc_b200 as hepmc
from hepmc_b200 import densities, proposals, DefaultMetropolis

hepmc.seed(1234)
target = densities.Banana(2, bananicity=0.1)
sampler = DefaultMetropolis(2, target, proposals.Gaussian(2, cov=10 * np.eye(2)))

chain = sampler.sample(10_000, [0., 10.])
print("one chain :", chain.data.shape, "acceptance %.3f" % chain.accept_ratio)

ensemble = sampler.sample(20_000, np.tile([0., 10.], (100_000, 1)), store_chain=False)
final = ensemble.data[:, 0]
print("ensemble  : 100000 chains x 20000 steps, acceptance %.3f" % (ensemble.accepted.mean() / 20000))
print("            var(x0) = %.1f (exact 100), var(x1) = %.1f (exact 201)" % (final[:, 0].var(), final[:, 1].var()))
