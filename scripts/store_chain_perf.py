"""RWM / HMC ensembles with store_chain=True (full chains returned, the reference's behaviour)
against final-states-only (development script)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
H.set_outputs("torch")

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts)

met = H.DefaultMetropolis(2, H.densities.Banana(2), H.proposals.Gaussian(2, cov=10 * np.eye(2)))
for C, T in ((100_000, 1001), (1_000_000, 501), (1000, 100_001)):
    x0 = torch.tensor([[0., 10.]], device="cuda", dtype=torch.float64).repeat(C, 1)
    for store in (False, True):
        t = timed(lambda: met.sample(T, x0, store_chain=store))
        print("RWM banana-2D C %7d T %6d store_chain=%-5s: %.3e chain-steps/s (%.0f GB/s written)"
              % (C, T, store, C * (T - 1) / t, (C * T * 16 / t / 1e9) if store else 0))
tgt = H.densities.Camel(8)
hmc = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8), 20, 0.01)
for C, T in ((100_000, 51),):
    x0 = torch.full((C, 8), 1 / 3, device="cuda", dtype=torch.float64)
    for store in (False, True):
        t = timed(lambda: hmc.sample(T, x0, store_chain=store))
        print("HMC camel-8D exact gradient C %7d T %4d store_chain=%-5s: %.3e chain-steps/s" % (C, T, store, C * (T - 1) / t))
