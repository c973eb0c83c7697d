"""Hand-off trace of the tensor-core ELM pipeline (library built with make EXTRA=-DHEPMC_TC_TRACE)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
from hepmc_b200 import _lib
H.set_outputs("torch")
lib = _lib.load()
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 512
rs = np.random.default_rng(0)
x0h = torch.full((148 * 256, 8), 1 / 3, device="cuda", dtype=torch.float64)
basis = H.surrogate.GaussianBasis(8)
params = (rs.random((nodes, 8)), 0.01 + 0.99 * rs.random(nodes), None)
w = 0.01 * rs.standard_normal(nodes)
tgt = H.densities.Camel(8)
H.surrogate.attach_surrogate_gradient(tgt, basis, params, 0.0, w)
upd = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8), 20, 0.01, precision="tc")
assert lib.hepmc_tc_debug(mode) == 0, "build the library with make EXTRA=-DHEPMC_TC_DEBUG"
upd.sample(2, x0h, store_chain=False); torch.cuda.synchronize()
buf = torch.zeros(3 * 8192, dtype=torch.int64, device="cuda")
lib.hepmc_tc_trace.argtypes = [ctypes.c_void_p]
lib.hepmc_tc_trace(ctypes.c_void_p(buf.data_ptr()))
upd.sample(2, x0h, store_chain=False); torch.cuda.synchronize()
lib.hepmc_tc_trace(ctypes.c_void_p(0))
r = buf.cpu().numpy().reshape(8192, 3)
names = {1: "own wait S_full", 2: "own got  S_full", 3: "own arrive P_ready", 4: "own epilogue done/A1 start", 5: "own arrived A_ready",
         6: "own got D2_full", 10: "iss wait A_ready", 11: "iss got  A_ready", 12: "iss G1+commit", 13: "iss got  P_ready",
         14: "iss G2 issued", 15: "iss commit D2"}
ev = [(int(c), int(t), int(k), "own") for k, t, c in r[:4096] if k] + [(int(c), int(t), int(k), "iss") for k, t, c in r[4096:] if k]
ev.sort()
t0 = ev[0][0]
# print evaluations 3..4 (steady state)
starts = [i for i, e in enumerate(ev) if e[2] == 4]
lo, hi = starts[3], starts[5]
prev = ev[lo][0]
for c, t, k, who in ev[lo:hi]:
    print("%8d (+%5d) %-28s tile %2d" % (c - ev[lo][0], c - prev, names[k], t))
    prev = c
