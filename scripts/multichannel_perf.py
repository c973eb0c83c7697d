"""MultiChannelMC throughput (development script)."""
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

for ndim, n in ((2, 20_000_000), (8, 20_000_000)):
    D = H.densities
    channels = H.MultiChannel([D.Gaussian(ndim, mu=1 / 3, scale=0.1), D.Gaussian(ndim, mu=2 / 3, scale=0.1), D.Uniform(ndim)])
    mc = H.MultiChannelMC(channels)
    tgt = D.Camel(ndim)
    for keep in (False, True):
        t = timed(lambda: mc(tgt, [], [n, n], [], keep_samples=keep))
        r = mc(tgt, [], [n, n], [], keep_samples=False)
        print("MultiChannelMC ndim %d, 3 channels, 2 x %.0e points keep=%-5s: %.3e points/s  integral %.5f +- %.5f"
              % (ndim, n, keep, 2 * n / t, r.integral, r.integral_err))

# where does keep_samples=True spend its time?
D = H.densities
ndim, n = 8, 20_000_000
channels = H.MultiChannel([D.Gaussian(ndim, mu=1 / 3, scale=0.1), D.Gaussian(ndim, mu=2 / 3, scale=0.1), D.Uniform(ndim)])
blk = D.Camel(ndim)._block()
t0 = timed(lambda: channels._launch(blk, n, (False, False, False, False)))
t1 = timed(lambda: channels._launch(blk, n, (True, True, True, True)))
x, f, p, ch, *_ = channels._launch(blk, n, (True, True, True, True))
t2 = timed(lambda: channels._grouped(x, ch, f, p))
parts = [torch.nonzero(ch == c).reshape(-1) for c in range(3)]
order = torch.cat(parts)
t3 = timed(lambda: [torch.nonzero(ch == c).reshape(-1) for c in range(3)])
t4 = timed(lambda: channels._grouped(x, ch, f, p))
print("launch no outputs %.2f ms, with outputs %.2f ms, grouped %.2f ms (nonzero passes %.2f ms, gathers %.2f ms)"
      % (t0 * 1e3, t1 * 1e3, t2 * 1e3, t3 * 1e3, t4 * 1e3))
