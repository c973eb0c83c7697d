"""Where does the per-iteration time outside the sample kernel go?  Back-to-back launches on one stream,
CUDA events around 20 repetitions: sample only / reduce only / update only / whole step."""
import ctypes, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
from hepmc_b200 import _lib, rng, parallel
H.set_outputs("torch")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10 ** 9
ndim, K = 16, 15
cols = ndim * K
chunk = _lib.default_chunk(n)
ws_emul = int(sys.argv[2]) if len(sys.argv) > 2 else 1     # shape of rank 0 of a ws_emul-GPU run, on one GPU
first, n_local, begins = parallel.local_span(n, chunk, 0, ws_emul) if ws_emul > 1 else parallel.local_span(n, chunk)
nslots = len(begins) - 1
nc = begins[-1]
dev = torch.device("cuda")
sp = torch.zeros((nc, 3), dtype=torch.float64, device=dev)
hp = torch.zeros((nc, cols), dtype=torch.float64, device=dev)
packed = torch.zeros((nslots, 3 + cols), dtype=torch.float64, device=dev)
scratch = torch.zeros(int(_lib.load().hepmc_reduce_slots_scratch_bytes(nslots, cols)) // 8 + 1, dtype=torch.float64, device=dev)
ev = torch.zeros((64, 2), dtype=torch.float64, device=dev)
camel = H.densities.Camel(ndim)
blk = camel._block()
mc = H.VegasMC(ndim, divisions=K)
mc(camel, 10 ** 7, 3, keep_samples=False)
grid = mc.volumes.device_grid()
bc = _lib.i64_array(begins)
st = _lib.stream_ptr()
def sample(i):
    _lib.call("hepmc_vegas_sample", ctypes.byref(blk), _lib.ptr(grid), ndim, K, n_local, first, chunk, 1234, 100 + i,
              rng.mode_code(), None, None, None, None, None, None, _lib.ptr(sp), _lib.ptr(hp), st)
def reduce(i):
    _lib.call("hepmc_reduce_slots", _lib.ptr(sp), _lib.ptr(hp), bc, nslots, cols, _lib.ptr(packed), _lib.ptr(scratch), st)
def update(i):
    _lib.call("hepmc_vegas_update_grid_packed", _lib.ptr(packed), nslots, _lib.ptr(grid), ndim, K, 3.0, i % 60, _lib.ptr(ev), st)
def timed(fns, reps=20):
    for f in fns: f(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        for f in fns: f(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
out = {"sample_us": timed([sample]), "reduce_us": timed([reduce]), "update_us": timed([update]),
       "reduce_update_us": timed([reduce, update]), "step_us": timed([sample, reduce, update])}
out["step_minus_sample_us"] = out["step_us"] - out["sample_us"]
print(json.dumps(out, indent=1))
