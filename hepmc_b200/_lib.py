"""ctypes binding of libhepmc_b200.so (the C ABI in include/hepmc_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is
visible when a compute entry point is called, a RuntimeError is raised.
PyTorch is used only for device memory, streams and torch.distributed.
"""
import ctypes
import os

import numpy as np
import torch

MAX_DIM = 32
MAX_CHAIN_DIM = 16
MAX_CHANNELS = 16
CHANNEL_DIM = 16
CHANNEL_ROW = 8 + 3 * CHANNEL_DIM
N_SLOTS = 8

CAMEL, UCAMEL, BANANA, GAUSSIAN, UNIFORM, NONE = range(6)
PDF, POT, POT_GRADIENT, PDF_GRADIENT = range(4)
GRAD_EXACT, GRAD_ELM_GAUSS, GRAD_ELM_TRIG = range(3)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HEPMC_B200_LIB", os.path.join(_HERE, "libhepmc_b200.so"))  # override: development A/B builds


class DensityBlock(ctypes.Structure):
    """hepmc_density_t"""
    _fields_ = [("kind", ctypes.c_int32), ("ndim", ctypes.c_int32),
                ("a", ctypes.c_double * MAX_DIM), ("b", ctypes.c_double * MAX_DIM),
                ("c", ctypes.c_double * MAX_DIM),
                ("s0", ctypes.c_double), ("s1", ctypes.c_double),
                ("s2", ctypes.c_double), ("s3", ctypes.c_double)]


class ElmBlock(ctypes.Structure):
    """hepmc_elm_t"""
    _fields_ = [("mode", ctypes.c_int32), ("nodes", ctypes.c_int32),
                ("p0_dev", ctypes.c_void_p), ("p1_dev", ctypes.c_void_p),
                ("out_w_dev", ctypes.c_void_p), ("out_bias", ctypes.c_double)]


_vp, _i64, _u64, _u32, _int, _dbl = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64,
                                     ctypes.c_uint32, ctypes.c_int, ctypes.c_double)
_dens_p = ctypes.POINTER(DensityBlock)
_elm_p = ctypes.POINTER(ElmBlock)

# name -> (restype, argtypes); mirrors include/hepmc_b200.h one to one
SIGNATURES = {
    "hepmc_abi_version": (_int, []),
    "hepmc_last_error": (ctypes.c_char_p, []),
    "hepmc_device_info": (_int, [_int, ctypes.POINTER(_i64)]),
    "hepmc_fp64_probe": (_int, [_i64, ctypes.POINTER(_dbl), ctypes.POINTER(_dbl), _vp]),
    "hepmc_mufu_probe": (_int, [_i64, ctypes.POINTER(_dbl), ctypes.POINTER(_dbl), _vp]),
    "hepmc_density_eval": (_int, [_dens_p, _int, _vp, _i64, _vp, _vp]),
    "hepmc_default_chunk": (_i64, [_i64]),
    "hepmc_plain_mc_sample": (_int, [_dens_p, _i64, _i64, _i64, _u64, _u32, _int, _vp, _vp, _vp, _vp, _vp]),
    "hepmc_reduce_stats": (_int, [_vp, ctypes.POINTER(_i64), _int, _vp, _vp]),
    "hepmc_stats_partials": (_int, [_vp, _vp, _dbl, _i64, _i64, _vp, _vp]),
    "hepmc_vegas_sample": (_int, [_dens_p, _vp, _int, _int, _i64, _i64, _i64, _u64, _u32, _int,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hepmc_vegas_accumulate": (_int, [_vp, _vp, _vp, _int, _int, _i64, _i64, _vp, _vp, _vp]),
    "hepmc_reduce_hist": (_int, [_vp, ctypes.POINTER(_i64), _int, _int, _vp, _vp]),
    "hepmc_vegas_update_grid": (_int, [_vp, _vp, _int, _vp, _int, _int, _dbl, _int, _vp, _vp]),
    "hepmc_reduce_slots_scratch_bytes": (_i64, [_int, _int]),
    "hepmc_reduce_slots": (_int, [_vp, _vp, ctypes.POINTER(_i64), _int, _int, _vp, _vp, _vp]),
    "hepmc_vegas_update_grid_packed": (_int, [_vp, _int, _vp, _int, _int, _dbl, _int, _vp, _vp]),
    "hepmc_slot_exchange_bytes": (_i64, [_int]),
    "hepmc_reduce_slots_p2p": (_int, [_vp, _vp, ctypes.POINTER(_i64), _int, _int, _vp, _vp, ctypes.POINTER(_i64), _int, _int,
                                      _u64, _vp]),
    "hepmc_vegas_update_grid_p2p": (_int, [_vp, _int, _u64, _vp, _int, _int, _dbl, _int, _vp, _vp, _vp]),
    "hepmc_vegas_workspace_bytes": (_i64, [_int, _int, _i64]),
    "hepmc_vegas_integrate": (_int, [_dens_p, _vp, _int, _int, _i64, _int, _dbl, _u64, _u32, _int,
                                     _vp, _vp, _i64, _vp]),
    "hepmc_grid_pdf": (_int, [_vp, _int, _int, _vp, _i64, _vp, _vp]),
    "hepmc_multichannel_sample": (_int, [_dens_p, _vp, _int, _int, _i64, _i64, _i64, _u64, _u32,
                                         _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hepmc_multichannel_accumulate": (_int, [_vp, _vp, _vp, _int, _i64, _i64, _vp, _vp, _vp]),
    "hepmc_rambo_map": (_int, [_int, _dbl, _i64, _i64, _u64, _u32, _int, _vp, _vp, _vp]),
    "hepmc_rambo_diet_map": (_int, [_int, _dbl, _i64, _i64, _u64, _u32, _vp, _vp, _vp]),
    "hepmc_rambo_diet_inverse": (_int, [_int, _dbl, _i64, _vp, _vp, _vp]),
    "hepmc_accept_flags": (_int, [_vp, _vp, _dbl, _dbl, _i64, _i64, _u64, _u32, _vp, _vp]),
    "hepmc_rambo_pdf": (_dbl, [_int, _dbl]),
    "hepmc_rwm_chains": (_int, [_dens_p, _vp, _int, ctypes.POINTER(_dbl), _i64, _i64, _i64, _i64, _u64, _u32,
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hepmc_leapfrog": (_int, [_dens_p, _elm_p, _int, _int, ctypes.POINTER(_dbl), _dbl, _int, _i64, _vp, _vp, _vp, _vp, _vp]),
    "hepmc_hmc_chains": (_int, [_dens_p, _elm_p, _int, _int, _vp, ctypes.POINTER(_dbl), _dbl, _int,
                                _i64, _i64, _i64, _i64, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hepmc_mc3_chains": (_int, [_dens_p, _vp, _int, _dbl, _int, ctypes.POINTER(_dbl), _dbl, _int, _elm_p, _int, _int, _vp,
                                _i64, _i64, _i64, _i64, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hepmc_elm_eval": (_int, [_elm_p, _int, _int, _int, _vp, _i64, _vp, _vp]),
    "hepmc_gather_rows": (_int, [_vp, _vp, _i64, _int, _vp, _vp]),
    "hepmc_autocov": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _i64, _i64, _vp, _vp]),
    "hepmc_effective_sample_size": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp]),
    "hepmc_histogramdd": (_int, [_vp, _i64, _int, _vp, ctypes.POINTER(_int), _vp, _vp]),
    "hepmc_bin_points": (_int, [_int, _vp, ctypes.POINTER(_int), _vp, _i64, _int, _u64, _u32, _vp, _vp, _vp]),
    "hepmc_bin_chi2": (_int, [_vp, _vp, _vp, _i64, _int, _dbl, _dbl, _dbl, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "hepmc_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C hepmc_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("hepmc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")


def check(rc, what=""):
    if rc != 0:
        msg = load().hepmc_last_error().decode("utf-8", "replace")
        raise RuntimeError("hepmc_b200 %s failed (code %d): %s" % (what, rc, msg))


def call(name, *args):
    """Invoke a C-ABI function on the current torch CUDA stream and check its status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    check(rc, name)


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a (contiguous) CUDA tensor, or NULL for None."""
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def to_device(x, dtype=torch.float64):
    """numpy / list / scalar / tensor -> contiguous CUDA tensor (+ flag: came from host)."""
    require_cuda()
    if isinstance(x, torch.Tensor):
        host = not x.is_cuda
        t = x.to(device=device(), dtype=dtype).contiguous()
        return t, host
    arr = np.ascontiguousarray(np.asarray(x, dtype=np.float64 if dtype == torch.float64 else None))
    t = torch.from_numpy(arr)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(device(), non_blocking=False).contiguous(), True


def to_host_like(t, host):
    """Return a NumPy array when the caller handed us host data, else the CUDA tensor."""
    return t.cpu().numpy() if host else t


def dbl_array(values, n=None):
    values = np.asarray(values, dtype=np.float64).ravel()
    n = len(values) if n is None else n
    arr = (ctypes.c_double * n)()
    for i, v in enumerate(values):
        arr[i] = float(v)
    return arr


def i64_array(values):
    arr = (ctypes.c_int64 * len(values))()
    for i, v in enumerate(values):
        arr[i] = int(v)
    return arr


def default_chunk(n_total):
    return int(load().hepmc_default_chunk(int(n_total)))


def device_info(dev=None):
    require_cuda()
    out = (ctypes.c_int64 * 5)()
    check(load().hepmc_device_info(torch.cuda.current_device() if dev is None else dev, out), "device_info")
    return dict(sm_count=out[0], cc=(out[1], out[2]), max_smem=out[3], clock_khz=out[4])


def fp64_probe(iters=1 << 15):
    """Measured FP64 DFMA throughput in flop/s (the FP64-SIMT roofline denominator)."""
    require_cuda()
    ms, fl = ctypes.c_double(), ctypes.c_double()
    check(load().hepmc_fp64_probe(int(iters), ctypes.byref(ms), ctypes.byref(fl), stream_ptr()), "fp64_probe")
    return fl.value / (ms.value * 1e-3)


def mufu_probe(iters=1 << 15):
    """Measured MUFU.EX2 throughput in exponentials/s (roofline denominator of the f32 / tensor-core ELM kernels)."""
    require_cuda()
    ms, ops = ctypes.c_double(), ctypes.c_double()
    check(load().hepmc_mufu_probe(int(iters), ctypes.byref(ms), ctypes.byref(ops), stream_ptr()), "mufu_probe")
    return ops.value / (ms.value * 1e-3)
