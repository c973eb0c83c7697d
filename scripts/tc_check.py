import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
H.set_outputs("torch")
rs = np.random.default_rng(0)
for nodes in (100, 128, 300, 1024):
    basis = H.surrogate.GaussianBasis(8)
    params = (rs.random((nodes, 8)), 0.01 + 0.99 * rs.random(nodes), None)
    w = rs.standard_normal(nodes)
    for J in (1, 100, 1000, 5000):
        xs = torch.as_tensor(rs.random((J, 8)), device="cuda")
        g64 = basis.eval_gradient(params, 0.0, w, xs, precision=0)
        g32 = basis.eval_gradient(params, 0.0, w, xs, precision=1)
        gtc = basis.eval_gradient(params, 0.0, w, xs, precision=2)
        torch.cuda.synchronize()
        sc = g64.abs().max().item()
        print("nodes %4d J %5d: |g| max %.3g  f32 err max %.2e  tc err max %.2e  tc err rms %.2e (relative to max |g|)" % (
            nodes, J, sc, (g32 - g64).abs().max().item() / sc, (gtc - g64).abs().max().item() / sc,
            ((gtc - g64) ** 2).mean().sqrt().item() / sc))
# timing
nodes = 1024
basis = H.surrogate.GaussianBasis(8)
params = (rs.random((nodes, 8)), 0.01 + 0.99 * rs.random(nodes), None)
w = rs.standard_normal(nodes)
xs = torch.as_tensor(rs.random((100000, 8)), device="cuda")
for prec in (0, 1, 2):
    basis.eval_gradient(params, 0.0, w, xs, precision=prec)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        basis.eval_gradient(params, 0.0, w, xs, precision=prec)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 5 * 1e-3
    print("precision %d: %.3f ms per 1e5 x 1024 gradient -> %.3g pair-evals/s" % (prec, t * 1e3, 1e5 * 1024 / t))
