"""Plain Monte Carlo integration over the unit hypercube (hepmc/core/integration/
integration.py).  `PlainMC(ndim)(fn, eval_count)` keeps the reference signature; the
uniform draw, the integrand and the mean / variance reduction are one fused kernel when
`fn` is a built-in density."""
import ctypes
import math

import torch

from ..sampling import Sample
from ..density import as_builtin
from ... import _lib, rng, parallel, config
from . import engine


class IntegrationSample(Sample):

    def __init__(self, **kwargs):
        self.function_values = None
        self.integral = None
        self.integral_err = None
        super().__init__(**kwargs)


class PlainMC(object):
    """integral = mean f(U), err = sqrt(var0(f)/N), U ~ uniform on [0,1]^ndim."""

    def __init__(self, ndim=1, name="MC Plain"):
        self.method_name = name
        self.ndim = ndim

    def __call__(self, fn, eval_count, keep_samples=True, replay=None):
        """:param fn: integrand taking ndim arrays (a built-in density is fused).
        :param eval_count: number of points (whole job, all GPUs).
        :param keep_samples: store data / function_values like the reference does
            (this rank's shard under torch.distributed).
        :param replay: (eval_count, ndim) uniforms to consume instead of the native
            generator, e.g. the reference's own np.random.random draws (single process)."""
        _lib.require_cuda()
        n_total = int(eval_count)
        dev = _lib.device()
        chunk = _lib.default_chunk(n_total)
        first, n_local, begins = parallel.local_span(n_total, chunk)
        dens = as_builtin(fn)
        u_replay = None
        if replay is not None:
            if parallel.world()[1] != 1:
                raise RuntimeError("replay mode is single-process")
            u_replay, _ = _lib.to_device(replay)
            if tuple(u_replay.shape) != (n_total, self.ndim):
                raise ValueError("replay draws must have shape (eval_count, ndim)")
        fused = dens is not None
        if fused and dens.ndim != self.ndim:
            raise RuntimeWarning("Mismatching dimensions.")
        need_data = keep_samples or not fused
        data = torch.empty((n_local, self.ndim), dtype=torch.float64, device=dev) if need_data else None
        fvals = torch.empty(n_local, dtype=torch.float64, device=dev) if (keep_samples and fused) else None
        nc_local = begins[-1]
        partials = torch.zeros((max(nc_local, 1), 3), dtype=torch.float64, device=dev)
        blk = dens._block() if fused else engine.none_block(self.ndim)
        sid = rng.next_stream()
        if n_local:
            _lib.call("hepmc_plain_mc_sample", ctypes.byref(blk), n_local, first, chunk, rng.get_seed(), sid,
                      rng.mode_code(),
                      _lib.ptr(u_replay), _lib.ptr(data), _lib.ptr(fvals), _lib.ptr(partials), _lib.stream_ptr())
        if not fused:
            fvals = engine.call_integrand(fn, [data[:, d] for d in range(self.ndim)])
            if n_local:
                _lib.call("hepmc_stats_partials", _lib.ptr(fvals), None, 1.0, n_local, chunk, _lib.ptr(partials),
                          _lib.stream_ptr())
        n, mean, m2 = engine.merged_stats(partials, begins)
        sample = IntegrationSample()
        if keep_samples:
            sample.data = config.deliver(data)
            sample.function_values = config.deliver(fvals)
        sample.integral = mean
        sample.integral_err = math.sqrt(m2 / n / n) if n else float("nan")
        return sample
