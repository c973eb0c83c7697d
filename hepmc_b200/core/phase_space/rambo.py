"""RAMBO massless phase space (hepmc/core/phase_space/rambo.py:21-69,162-173).

`Rambo(e_cm, nparticles).map(xs)`: (N, 4n) unit-cube points -> (N, 4n) momenta
[E,px,py,pz] x n summing to (e_cm,0,0,0); `pdf` is the constant 1/Phi_n (a Python
scalar, like the reference).  `generate(n)` is the native-RNG path (uniforms never leave
the registers)."""
import math

import numpy as np
import torch

from ... import _lib, rng, parallel, config


class PhaseSpaceMapping(object):
    def __init__(self, ndim):
        self.ndim = ndim

    def pdf(self, xs):
        raise NotImplementedError

    def map(self, xs):
        raise NotImplementedError


class Rambo(PhaseSpaceMapping):

    def __init__(self, e_cm, nparticles):
        self.e_cm = e_cm
        self.nparticles = nparticles
        super().__init__(nparticles * 4)

    def pdf(self, xs):
        n = self.nparticles
        if n is None:
            n = xs.shape[1] // 4
        vol = ((math.pi / 2.) ** (n - 1) * self.e_cm ** (2 * n - 4) /
               (math.gamma(n) * math.gamma(n - 1)))
        return 1 / vol

    def map(self, xs):
        """Map given unit-cube points (the reference's own draws in replay tests)."""
        x_dev, host = _lib.to_device(xs)
        if x_dev.dim() != 2 or x_dev.shape[1] != self.ndim:
            raise RuntimeWarning("Unexpected dimension of entries in array.")
        out = torch.empty_like(x_dev)
        if x_dev.shape[0]:
            _lib.call("hepmc_rambo_map", int(self.nparticles), float(self.e_cm), x_dev.shape[0], 0, 0, 0, 0,
                      _lib.ptr(x_dev), _lib.ptr(out), _lib.stream_ptr())
        return _lib.to_host_like(out, host)

    def generate(self, count, first_index=0, stream_id=None, out=None):
        """`count` phase-space points from the native generator: point i uses Philox
        counter first_index + i.  Returns a (count, 4n) CUDA tensor (written into `out`
        when given)."""
        dev = _lib.device()
        if out is None:
            out = torch.empty((count, self.ndim), dtype=torch.float64, device=dev)
        sid = rng.next_stream() if stream_id is None else stream_id
        if count:
            _lib.call("hepmc_rambo_map", int(self.nparticles), float(self.e_cm), int(count), int(first_index),
                      rng.get_seed(), sid, rng.mode_code(), None, _lib.ptr(out), _lib.stream_ptr())
        return out


class RamboOnDiet(PhaseSpaceMapping):
    """Invertible massless phase-space map with 3n-4 variables (rambo.py:72-159)."""

    def __init__(self, e_cm, nparticles):
        self.e_cm = e_cm
        self.nparticles = nparticles
        super().__init__(nparticles * 3 - 4)

    def map(self, xs):
        from ..util import interpret_array
        x_dev, host = _lib.to_device(interpret_array(xs, self.ndim))
        out = torch.empty((x_dev.shape[0], 4 * self.nparticles), dtype=torch.float64, device=x_dev.device)
        if x_dev.shape[0]:
            _lib.call("hepmc_rambo_diet_map", int(self.nparticles), float(self.e_cm), x_dev.shape[0], 0, 0, 0,
                      _lib.ptr(x_dev), _lib.ptr(out), _lib.stream_ptr())
        return _lib.to_host_like(out, host)

    def generate(self, count, first_index=0, stream_id=None):
        """`count` points from the native generator -> (count, 4n) CUDA tensor."""
        out = torch.empty((count, 4 * self.nparticles), dtype=torch.float64, device=_lib.device())
        sid = rng.next_stream() if stream_id is None else stream_id
        if count:
            _lib.call("hepmc_rambo_diet_map", int(self.nparticles), float(self.e_cm), int(count), int(first_index),
                      rng.get_seed(), sid, None, _lib.ptr(out), _lib.stream_ptr())
        return out

    def map_inverse(self, p):
        p_dev, host = _lib.to_device(p)
        p_dev = p_dev.reshape(p_dev.shape[0], -1).contiguous()
        out = torch.empty((p_dev.shape[0], self.ndim), dtype=torch.float64, device=p_dev.device)
        if p_dev.shape[0]:
            _lib.call("hepmc_rambo_diet_inverse", int(self.nparticles), float(self.e_cm), p_dev.shape[0],
                      _lib.ptr(p_dev), _lib.ptr(out), _lib.stream_ptr())
        return _lib.to_host_like(out, host)

    def pdf(self, xs):
        n = self.nparticles
        if n is None:
            n = xs.shape[1] // 4
        vol = ((math.pi / 2.) ** (n - 1) * self.e_cm ** (2 * n - 4) / (math.gamma(n) * math.gamma(n - 1)))
        return 1 / vol
