import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hepmc_b200 as H
H.set_outputs("torch")
m = H.phase_space.Rambo(100., 4)
n = 1 << 26
buf = torch.empty((n, 16), dtype=torch.float64, device="cuda")
for i in range(3):
    m.generate(n, out=buf)
torch.cuda.synchronize()
print("ok")
