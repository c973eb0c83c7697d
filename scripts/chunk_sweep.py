"""Headline sample kernel at several chunk sizes (per-CTA fixed cost vs wave quantisation); development script."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
from hepmc_b200 import _lib
H.set_outputs("torch")
NDIM, K = 16, 15
camel = H.densities.Camel(NDIM)
mc = H.VegasMC(NDIM, divisions=K)
grid = mc.volumes.device_grid()
blk = camel._block()
for n in (10**9, 10**9 // 8):
    for chunk in (4096, 8192, 16384, 32768, 65536):
        nc = (n + chunk - 1) // chunk
        sp = torch.zeros((nc, 3), dtype=torch.float64, device="cuda")
        hp = torch.zeros((nc, NDIM * K), dtype=torch.float64, device="cuda")
        ts = []
        for i in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.call("hepmc_vegas_sample", ctypes.byref(blk), _lib.ptr(grid), NDIM, K, n, 0, chunk, 1234, 100 + i,
                      H.rng.mode_code(), None, None, None, None, None, None, _lib.ptr(sp), _lib.ptr(hp), _lib.stream_ptr())
            b.record(); torch.cuda.synchronize()
            if i >= 2: ts.append(a.elapsed_time(b))
        print("n %.3g chunk %6d: %4d CTAs = %6.2f waves  %.3f ms  %.3e points/s" % (n, chunk, nc, nc / 444, np.mean(ts), n / np.mean(ts) * 1e3))
