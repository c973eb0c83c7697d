"""Density seen through a phase-space mapping (mirror of hepmc/core/phase_space/mapping.py):
pdf(xs) = density.pdf(mapping.map(xs)) / mapping.pdf(xs) [/ norm] on the open unit cube, 0 outside.
Composition of device calls; the elementwise glue runs in torch on the device."""
import math

import numpy as np
import torch

from ..density import Density
from ..util import interpret_array
from ... import _lib


class MappedDensity(Density):

    def __init__(self, density, mapping, norm=None):
        super().__init__(mapping.ndim)
        self.density = density
        self.mapping = mapping
        self.norm = norm

    def _bounded(self, xs, fn, null_value=0.0):
        x_dev, host = _lib.to_device(interpret_array(xs, self.ndim))
        inside = ((x_dev > 0) & (x_dev < 1)).all(dim=1)
        res = torch.full((x_dev.shape[0],), float(null_value), dtype=torch.float64, device=x_dev.device)
        if bool(inside.any()):
            res[inside] = fn(x_dev[inside].contiguous())
        return _lib.to_host_like(res, host)

    def pdf(self, xs):
        def fn(x):
            ps = self.mapping.map(x)
            pdf = torch.as_tensor(self.density.pdf(ps), dtype=torch.float64, device=x.device) / self.mapping.pdf(x)
            return pdf / self.norm if self.norm is not None else pdf
        return self._bounded(xs, fn)

    def pot(self, xs):
        def fn(x):
            ps = self.mapping.map(x)
            pot = torch.as_tensor(self.density.pot(ps), dtype=torch.float64, device=x.device)
            return pot - math.log(self.norm) - math.log(self.mapping.pdf(x))
        return self._bounded(xs, fn)

    def pdf_gradient(self, xs):
        raise NotImplementedError  # mapping pdf gradient generally not known

    def pot_gradient(self, xs):
        raise NotImplementedError
