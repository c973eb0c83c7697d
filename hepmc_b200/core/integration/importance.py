"""Importance-sampling Monte Carlo integration (hepmc/core/integration/importance.py:5-51).

`ImportanceMC(dist)(fn, eval_count)`: xs = dist.rvs(N), ys = fn(*xs.T), w = dist.pdf(xs),
integral = mean(ys/w), err = sqrt(var0(ys/w)/N).  `dist` is any Distribution -- the built-in
ones (densities.Rambo, GridVolumes, Uniform, Gaussian) sample on the device; the ratio,
mean and variance always reduce on the device (hepmc_stats_partials / hepmc_reduce_stats).
"""
import math

import numpy as np
import torch

from .integration import IntegrationSample
from ..density import as_builtin
from ... import _lib, config
from . import engine


class ImportanceMC(object):

    def __init__(self, dist, name="MC Importance"):
        self.method_name = name
        self.dist = dist
        self.ndim = dist.ndim

    def __call__(self, fn, eval_count, keep_samples=True):
        _lib.require_cuda()
        xs = self.dist.rvs(int(eval_count))
        xs_dev, _ = _lib.to_device(xs)
        dens = as_builtin(fn)
        if dens is not None:
            ys = dens.pdf(xs_dev)
        else:
            ys = engine.call_integrand(fn, [xs_dev[:, d] for d in range(xs_dev.shape[1])])
        weights = self.dist.pdf(xs_dev if isinstance(xs, torch.Tensor) else xs)
        if isinstance(weights, np.ndarray):
            weights = torch.as_tensor(weights, dtype=torch.float64, device=xs_dev.device)
        n, mean, m2 = engine.stats_of_values(ys, weights, n_total=int(eval_count) if xs_dev.shape[0] != int(eval_count) else None)
        sample = IntegrationSample()
        if keep_samples:
            sample.data = config.deliver(xs_dev)
            sample.function_values = config.deliver(ys)
            sample.weights = config.deliver(weights) if isinstance(weights, torch.Tensor) else weights
        sample.integral = mean
        sample.integral_err = math.sqrt(m2 / n / n) if n else float("nan")
        return sample
