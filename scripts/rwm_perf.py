"""banana-2D random-walk Metropolis throughput (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
H.set_outputs("torch")
C, T = 1_000_000, 2001
met = H.DefaultMetropolis(2, H.densities.Banana(2), H.proposals.Gaussian(2, cov=10 * np.eye(2)))
x0 = torch.tensor([[0., 10.]], device="cuda", dtype=torch.float64).repeat(C, 1)
met.sample(T, x0, store_chain=False); torch.cuda.synchronize()
ts = []
for i in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s = met.sample(T, x0, store_chain=False); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e-3)
print("rwm banana-2D: %.4g chain-steps/s, acceptance %.3f" % (C * (T - 1) / min(ts), float(s.accepted.double().mean()) / T))
