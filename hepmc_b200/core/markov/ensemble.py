"""Launch helpers for the chain-ensemble kernels (hepmc_rwm_chains, hepmc_hmc_chains).

An ensemble is `C` independent chains identified by GLOBAL chain indices
`first_chain .. first_chain + C - 1`; under torch.distributed each rank runs a contiguous
index range with no collective on the path (acceptance totals are summed on request).
"""
import ctypes

import numpy as np
import torch

from ... import _lib, rng


def gaussian_rvs(mean, scale, n):
    """(n, ndim) draws mean + scale * z, z standard normal from the native generator."""
    from ..integration.engine import uniform_rvs
    ndim = len(mean)
    npair = (ndim + 1) // 2
    u = uniform_rvs(n, 2 * npair)
    r = torch.sqrt(-2.0 * torch.log(u[:, 0::2]))
    ang = 2.0 * np.pi * u[:, 1::2]
    z = torch.stack((r * torch.cos(ang), r * torch.sin(ang)), dim=2).reshape(n, 2 * npair)[:, :ndim]
    dev = z.device
    return (torch.as_tensor(np.asarray(mean, dtype=np.float64), device=dev) +
            torch.as_tensor(np.asarray(scale, dtype=np.float64), device=dev) * z).contiguous()


def prepare_initial(initial, ndim):
    """initial -> ((C, ndim) CUDA tensor, is_ensemble).  A 1-D `initial` of length ndim is a
    single chain (the reference's only mode, base.py:56-59); a 2-D (C, ndim) array starts C
    independent chains (the signature-preserving ensemble extension)."""
    if isinstance(initial, torch.Tensor):
        arr = initial
        nd = arr.dim()
    else:
        arr = np.atleast_1d(np.asarray(initial, dtype=np.float64))
        nd = arr.ndim
    if nd == 1:
        if arr.shape[0] != ndim:
            raise ValueError('initial must have dimension ' + str(ndim))
        x0, _ = _lib.to_device(arr.reshape(1, ndim))
        return x0, False
    if nd == 2:
        if arr.shape[1] != ndim:
            raise ValueError('initial must have dimension ' + str(ndim))
        x0, _ = _lib.to_device(arr)
        return x0, True
    raise ValueError('initial must be (ndim,) or (n_chains, ndim)')


def alloc_outputs(n_chains, ndim, sample_size, thin, store_chain, dev):
    n_kept = (sample_size + thin - 1) // thin
    chain = torch.empty((n_chains, n_kept, ndim), dtype=torch.float64, device=dev) if store_chain else None
    last = torch.empty((n_chains, ndim), dtype=torch.float64, device=dev)
    accepted = torch.zeros(n_chains, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    return chain, last, accepted, total


def run_rwm(blk, x0, proposal_kind, prop_param, sample_size, thin=1, store_chain=True, first_chain=0,
            stream_id=None, replay=None):
    dev = x0.device
    n_chains, ndim = x0.shape
    chain, last, accepted, total = alloc_outputs(n_chains, ndim, sample_size, thin, store_chain, dev)
    z = u = None
    if replay is not None:
        z, _ = _lib.to_device(replay[0])
        u, _ = _lib.to_device(replay[1])
        if tuple(z.shape) != (n_chains, sample_size - 1, ndim) or tuple(u.shape) != (n_chains, sample_size - 1):
            raise ValueError("replay tapes must be (C, T-1, ndim) and (C, T-1)")
    sid = rng.next_stream() if stream_id is None else stream_id
    _lib.call("hepmc_rwm_chains", ctypes.byref(blk), _lib.ptr(x0), int(proposal_kind), _lib.dbl_array(prop_param),
              n_chains, int(first_chain), int(sample_size), int(thin), rng.get_seed(), sid,
              _lib.ptr(z), _lib.ptr(u), _lib.ptr(chain), _lib.ptr(last), _lib.ptr(accepted), _lib.ptr(total),
              _lib.stream_ptr())
    return chain, last, accepted, total


def run_hmc(blk, elm_blk, precision, x0, mass_diag, step_size, steps, sample_size, thin=1, store_chain=True,
            first_chain=0, stream_id=None, replay=None):
    dev = x0.device
    n_chains, ndim = x0.shape
    chain, last, accepted, total = alloc_outputs(n_chains, ndim, sample_size, thin, store_chain, dev)
    p = u = None
    if replay is not None:
        p, _ = _lib.to_device(replay[0])
        u, _ = _lib.to_device(replay[1])
        if tuple(p.shape) != (n_chains, sample_size - 1, ndim) or tuple(u.shape) != (n_chains, sample_size - 1):
            raise ValueError("replay tapes must be (C, T-1, ndim) and (C, T-1)")
    sid = rng.next_stream() if stream_id is None else stream_id
    _lib.call("hepmc_hmc_chains", ctypes.byref(blk), ctypes.byref(elm_blk) if elm_blk is not None else None,
              int(precision), _lib.ptr(x0), _lib.dbl_array(mass_diag), float(step_size), int(steps),
              n_chains, int(first_chain), int(sample_size), int(thin), rng.get_seed(), sid,
              _lib.ptr(p), _lib.ptr(u), _lib.ptr(chain), _lib.ptr(last), _lib.ptr(accepted), _lib.ptr(total),
              _lib.stream_ptr())
    return chain, last, accepted, total
