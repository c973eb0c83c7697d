"""VEGAS / plain MC with keep_samples=True (the reference's default: points, values and weights are
returned) against the fused no-output path (development script)."""
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

for ndim, K, n in ((16, 15, 50_000_000), (2, 50, 100_000_000), (8, 15, 50_000_000)):
    camel = H.densities.Camel(ndim)
    for keep in (False, True):
        mc = H.VegasMC(ndim, divisions=K)
        t = timed(lambda: mc(camel, n, 2, keep_samples=keep))
        byt = 2 * n * (ndim + 2) * 8 if keep else 0
        print("VEGAS ndim %2d K %2d n %.0e keep=%-5s: %.3e points/s  (%.0f GB/s written)" % (ndim, K, n, keep, 2 * n / t, byt / t / 1e9))
    for keep in (False, True):
        t = timed(lambda: H.PlainMC(ndim)(camel, n, keep_samples=keep))
        byt = n * (ndim + 1) * 8 if keep else 0
        print("plain ndim %2d      n %.0e keep=%-5s: %.3e points/s  (%.0f GB/s written)" % (ndim, n, keep, n / t, byt / t / 1e9))
