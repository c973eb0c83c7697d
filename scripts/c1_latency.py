import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hepmc_b200 as H
H.set_outputs("torch")
camel2 = H.densities.Camel(2)
mc2 = H.VegasMC(2, divisions=50)
for n, it in ((100_000, 10), (10_000, 10), (1_000_000, 10)):
    mc2(camel2, n, it, keep_samples=False); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        mc2(camel2, n, it, keep_samples=False)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("VegasMC(2,50) n %8d x %d iterations: %.1f us per call, %.1f us per iteration" % (n, it, dt * 1e6, dt * 1e6 / it))
