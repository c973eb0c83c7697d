"""Development aid: the three RNG modes of the VEGAS / plain kernels -- oracle check at small n,
kernel-only timing at 1e9 points (camel-16D, K=15).  usage: python scripts/vegas_modes.py [n]"""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hepmc_b200 as H
from hepmc_b200 import _lib, rng, parallel
from oracle import hepmc_oracle as O

H.set_outputs("torch")
n_big = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 9
res = {"lib": os.environ.get("HEPMC_B200_LIB", "default")}
for mode in ("u52r10", "u32r10", "u32r7"):
    rng.set_mode(mode)
    # ---- oracle check: 4-D and 16-D, two iterations, native draws
    for ndim, K, n in ((4, 8, 3000), (16, 15, 5000), (5, 50, 2000)):
        H.seed(99)
        mc = H.VegasMC(ndim, divisions=K)
        s = mc(H.densities.Camel(ndim), n, 2, keep_samples=True)
        tapes = [O.philox_vegas_draws(99, j, np.arange(n), ndim, K, mode) for j in range(2)]
        r = O.vegas(O.camel_pdf, ndim, K, np.stack([t[0] for t in tapes]), np.stack([t[1] for t in tapes]))
        err = abs(s.integral - r["integral"]) / abs(r["integral"])
        gerr = np.max(np.abs(np.stack(mc.volumes.bounds) - r["bounds_history"][-1]))
        xerr = float(np.max(np.abs(s.data.cpu().numpy() - r["data"]))) if "data" in r else -1
        res["%s_check_%dd" % (mode, ndim)] = [err, gerr, xerr]
        H.seed(99)
        s2 = H.VegasMC(ndim, divisions=K)(H.densities.Camel(ndim), n, 2, keep_samples=False)
        res["%s_keep_vs_nokeep_%dd" % (mode, ndim)] = abs(s2.integral - s.integral)
    # ---- timing: sample kernel alone
    ndim, K = 16, 15
    chunk = _lib.default_chunk(n_big)
    first, n_local, begins = parallel.local_span(n_big, chunk)
    nc = max(begins[-1], 1)
    sp = torch.zeros((nc, 3), dtype=torch.float64, device="cuda")
    hp = torch.zeros((nc, ndim * K), dtype=torch.float64, device="cuda")
    camel = H.densities.Camel(ndim)
    blk = camel._block()
    mc = H.VegasMC(ndim, divisions=K)
    mc(camel, 10 ** 7, 3, keep_samples=False)        # an adapted grid
    grid = mc.volumes.device_grid()
    ts = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.call("hepmc_vegas_sample", ctypes.byref(blk), _lib.ptr(grid), ndim, K, n_local, first, chunk,
                  1234, 1000 + i, rng.mode_code(), None, None, None, None, None, None, _lib.ptr(sp), _lib.ptr(hp), _lib.stream_ptr())
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    res["%s_kernel_ms" % mode] = float(np.mean(ts))
    res["%s_pts_per_s" % mode] = n_local / (np.mean(ts) * 1e-3)
    H.seed(1234)
    r = mc(camel, n_big, 3, keep_samples=False)
    res["%s_integral" % mode] = [r.integral, r.integral_err]
    # plain MC
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    H.PlainMC(16)(camel, 10 ** 8, keep_samples=False)
    t0.record(); pr = H.PlainMC(16)(camel, n_big, keep_samples=False); t1.record(); torch.cuda.synchronize()
    res["%s_plain16_pts_per_s" % mode] = n_big / (t0.elapsed_time(t1) * 1e-3)
rng.set_mode("u52r10")
print(json.dumps(res, indent=1))
