"""Host orchestration of the integration kernels (launch order, workspaces, the slot
exchange).  Everything numerical happens in libhepmc_b200.so; this file only sequences
launches on the current CUDA stream and never synchronises except to read final scalars.
"""
import ctypes
import math

import numpy as np
import torch

from ... import _lib, rng, parallel
from ..density import as_builtin
from ..util import host_function
from ..util.helper import call_user_function


def none_block(ndim):
    blk = _lib.DensityBlock()
    blk.kind = _lib.NONE
    blk.ndim = ndim
    return blk


def uniform_rvs(n, ndim, stream_id=None):
    """(n, ndim) uniforms in (0,1) from the native generator (a CUDA tensor)."""
    dev = _lib.device()
    out = torch.empty((n, ndim), dtype=torch.float64, device=dev)
    if n == 0:
        return out
    sid = rng.next_stream() if stream_id is None else stream_id
    blk = none_block(ndim)
    _lib.call("hepmc_plain_mc_sample", ctypes.byref(blk), n, 0, _lib.default_chunk(n), rng.get_seed(), sid,
              rng.MODES["u52r10"], None, _lib.ptr(out), None, None, _lib.stream_ptr())
    return out


def call_integrand(fn, cols):
    """Evaluate a user integrand on the sample columns (CUDA tensors, or host arrays for a
    `host_function`) and return the values as a float64 CUDA tensor."""
    if not isinstance(fn, host_function) and len(cols) > 1 and cols[0].stride(0) != 1:
        # columns of a row-major (N, ndim) block: one transposing copy, then the integrand's
        # elementwise ops run on contiguous arrays (several times faster than on stride-ndim views)
        cols = list(torch.stack(cols, dim=0).unbind(0))
    vals = call_user_function(fn, cols, "the integrand")
    n = cols[0].shape[0]
    if isinstance(vals, torch.Tensor):
        vals = vals.to(device=cols[0].device, dtype=torch.float64)
    else:
        vals = torch.as_tensor(np.asarray(vals, dtype=np.float64), device=cols[0].device)
    if vals.dim() == 0:
        vals = vals.expand(n)
    return vals.reshape(-1).contiguous()


def merged_stats(partials, local_begins):
    """Chunk partial rows of this rank -> (count, mean, M2) of the whole call, identical on
    every rank and at every rank count.  One tiny D2H copy."""
    dev = partials.device
    n_loc = len(local_begins) - 1
    slot_rows = torch.zeros((n_loc, 3), dtype=torch.float64, device=dev)
    _lib.call("hepmc_reduce_stats", _lib.ptr(partials), _lib.i64_array(local_begins), n_loc,
              _lib.ptr(slot_rows), _lib.stream_ptr())
    slots = parallel.gather_slots(slot_rows)
    total = torch.empty(3, dtype=torch.float64, device=dev)
    _lib.call("hepmc_reduce_stats", _lib.ptr(slots), _lib.i64_array([0, slots.shape[0]]), 1,
              _lib.ptr(total), _lib.stream_ptr())
    n, mean, m2 = total.cpu().tolist()
    return n, mean, m2


def stats_of_values(values, weights=None, n_total=None):
    """(count, mean, M2) of values (or values/weights): importance.py:47-50.

    Under torch.distributed `values` must be this rank's `parallel.local_span(n_total, chunk)`
    shard, so that chunks and slots are functions of the global index only."""
    n_local = values.shape[0]
    rank, ws = parallel.world()
    if n_total is None:
        if ws != 1:
            raise RuntimeError("n_total is required under torch.distributed")
        n_total = n_local
    chunk = _lib.default_chunk(n_total)
    first, expect, begins = parallel.local_span(n_total, chunk)
    if expect != n_local:
        raise RuntimeError("values are not this rank's local_span shard (%d != %d)" % (n_local, expect))
    nc = begins[-1]
    partials = torch.zeros((max(nc, 1), 3), dtype=torch.float64, device=values.device)
    wt, wscalar = None, 1.0
    if weights is not None:
        if isinstance(weights, torch.Tensor) and weights.dim() >= 1 and weights.numel() == n_local and n_local > 1:
            wt = weights.to(device=values.device, dtype=torch.float64).contiguous()
        else:
            wscalar = float(weights)
    if n_local:
        _lib.call("hepmc_stats_partials", _lib.ptr(values), _lib.ptr(wt), wscalar, n_local, chunk,
                  _lib.ptr(partials), _lib.stream_ptr())
    return merged_stats(partials, begins)
