"""Phase-space integration with RAMBO as the importance-sampling distribution
(the pattern of the reference's examples/rambo.py)."""
import numpy as np
import torch
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from anywhere
import hepmc_b200 as hepmc
from hepmc_b200 import densities, ImportanceMC

hepmc.seed(1234)
hepmc.set_outputs("torch")                 # keep samples on the device
E_CM, N = 100., 4
rambo = densities.Rambo(N, E_CM)


def matrix_element(*momenta):              # 16 columns: (E, px, py, pz) of the four particles
    p = torch.stack(momenta, dim=1).view(-1, N, 4)
    s12 = (p[:, 0, 0] + p[:, 1, 0]) ** 2 - ((p[:, 0, 1:] + p[:, 1, 1:]) ** 2).sum(dim=1)
    return s12 / E_CM ** 2                  # a smooth observable: invariant mass^2 of the first pair


sample = ImportanceMC(rambo)(matrix_element, 5_000_000)
volume = 1.0 / rambo.pdf(None)
print("phase-space volume   = %.6e" % volume)
print("<s12 / s>            = %.5f +- %.5f" % (sample.integral / volume, sample.integral_err / volume))
print("momentum conservation: max |sum p - (E,0,0,0)| = %.2e" %
      float((sample.data.view(-1, N, 4).sum(dim=1) - torch.tensor([E_CM, 0, 0, 0], device="cuda")).abs().max()))
