import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hepmc_b200 as H
H.set_outputs("torch")
camel = H.densities.Camel(16)
n = 50_000_000
for _ in range(2):
    H.PlainMC(16)(camel, n, keep_samples=True)
    H.VegasMC(16, divisions=15)(camel, n, 1, keep_samples=True)
torch.cuda.synchronize()
print("done")
