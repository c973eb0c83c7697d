"""RAMBO 2->4 throughput in the three RNG modes (development aid)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hepmc_b200 as H
H.set_outputs("torch")
m = H.phase_space.Rambo(100., 4)
n = 1 << 27
buf = torch.empty((n, 16), dtype=torch.float64, device="cuda")
out = {}
for mode in ("u52r10", "u32r10", "u32r7"):
    H.rng.set_mode(mode)
    m.generate(n, out=buf)
    torch.cuda.synchronize()
    ts = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); m.generate(n, out=buf); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sum(ts) / len(ts) * 1e-3
    out[mode] = {"pts_per_s": n / t, "GBs": n * 128 / t / 1e9}
print(json.dumps(out, indent=1))
