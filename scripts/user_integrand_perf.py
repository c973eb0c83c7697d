"""Two-kernel path: user integrand written with torch ops on CUDA tensors (development script)."""
import sys, os, math
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

def make_fn(ndim):
    norm = (2 * math.pi * 0.005) ** (-ndim / 2)
    def fn(*xs):
        ra = sum((x - 1 / 3) ** 2 for x in xs)
        rb = sum((x - 2 / 3) ** 2 for x in xs)
        return 0.5 * norm * (torch.exp(-ra / 0.01) + torch.exp(-rb / 0.01))
    return fn

for ndim, K, n in ((2, 50, 20_000_000), (16, 15, 10_000_000)):
    fn = make_fn(ndim)
    for keep in (False, True):
        mc = H.VegasMC(ndim, divisions=K)
        t = timed(lambda: mc(fn, n, 2, keep_samples=keep))
        r = mc(fn, n, 3, keep_samples=False)
        print("VEGAS user fn ndim %2d n %.0e keep=%-5s: %.3e points/s   integral %.4f +- %.4f" % (ndim, n, keep, 2 * n / t, r.integral, r.integral_err))
    t = timed(lambda: H.PlainMC(ndim)(fn, n, keep_samples=False))
    print("plain user fn ndim %2d n %.0e            : %.3e points/s" % (ndim, n, n / t))
    xs = [torch.rand(n, device="cuda", dtype=torch.float64) for _ in range(ndim)]
    t = timed(lambda: fn(*xs))
    print("   the torch integrand alone            : %.3e points/s" % (n / t))
