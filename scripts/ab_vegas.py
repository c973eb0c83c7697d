import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hepmc_b200 as H
H.set_outputs("torch")
n = 10 ** 9
mc = H.VegasMC(16, divisions=15)
H.seed(1234)
r = mc(H.densities.Camel(16), n, 2, keep_samples=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); r = mc(H.densities.Camel(16), n, 5, apriori=False, keep_samples=False); b.record(); torch.cuda.synchronize()
print(os.environ.get("HEPMC_B200_LIB", "default"), "%.4g pts/s" % (5 * n / (a.elapsed_time(b) * 1e-3)), r.integral, r.integral_err)
