"""Stratified Monte Carlo integration over a GridVolumes partition
(hepmc/core/integration/stratified.py:6-91).

integral = sum_bins vol * mean(f in bin),  err^2 = sum_bins var0(f in bin) * vol^2 / count.
All bins are sampled in ONE batch on the device (draw kernel -> map into the bins -> integrand:
the pdf kernel for a built-in density, the user's function on CUDA tensors otherwise); the per-bin
moments are taken bin-count group by bin-count group as (bins, count) matrices, so there is no
per-bin Python loop and no host arithmetic on the sample.
"""
import numpy as np
import torch

from .integration import IntegrationSample
from .stratified_volume import GridVolumes
from ..density import as_builtin
from ... import _lib, config
from . import engine


class StratifiedSample(IntegrationSample):
    def __init__(self, **kwargs):
        self.sub_sizes = []
        self.sub_weights = []
        super().__init__(**kwargs)

    @property
    def weights(self):
        """Per-point weight count / vol / total of the bin the point was drawn in."""
        if self.function_values is None or len(self.sub_sizes) == 0:
            return None
        w = np.repeat(np.asarray(self.sub_weights, dtype=np.float64), np.asarray(self.sub_sizes, dtype=np.int64))
        if isinstance(self.function_values, torch.Tensor):
            return torch.as_tensor(w, device=self.function_values.device)
        return w

    @weights.setter
    def weights(self, value):
        pass


class StratifiedMC(object):

    def __init__(self, volumes=None, name="MC Stratified"):
        self.method_name = name
        self.volumes = GridVolumes() if volumes is None else volumes
        self.ndim = self.volumes.ndim

    def get_interface_infer_multiple(self):
        """fn, eval_count -> self(fn, eval_count / total_base_count) (stratified.py:44-66)."""
        def interface(fn, eval_count):
            return self(fn, eval_count / self.volumes.total_base_count)
        interface.method_name = self.method_name
        return interface

    def __call__(self, fn, multiple, keep_samples=True, replay=None):
        """:param multiple: every bin gets multiple * its base count points (:68-91).
        :param replay: per-bin uniforms (count, ndim) in np.ndindex order, consumed instead of the
            native generator (the reference's np.random.rand draws)."""
        _lib.require_cuda()
        vol = self.volumes
        counts, lower, upper = vol._cells(multiple)
        xs, cell = vol._cell_points(counts, lower, upper, replay)
        dens = as_builtin(fn)
        if dens is not None:
            values = dens.pdf(xs)
        else:
            values = engine.call_integrand(fn, [xs[:, d] for d in range(self.ndim)])
        vols = torch.as_tensor(np.prod(upper - lower, axis=1), device=xs.device)
        cnt = torch.as_tensor(counts, device=xs.device)
        starts = torch.cumsum(cnt, 0) - cnt
        mean = torch.zeros(counts.shape[0], dtype=torch.float64, device=xs.device)
        var = torch.zeros_like(mean)
        for c in np.unique(counts):                       # bins with the same count form one (bins, count) matrix
            if c == 0:
                continue
            sel = torch.nonzero(cnt == int(c)).reshape(-1)
            rows = values[(starts[sel][:, None] + torch.arange(int(c), device=xs.device)[None, :])]
            mean[sel] = rows.mean(dim=1)
            var[sel] = rows.var(dim=1, unbiased=False)
        used = cnt > 0
        int_est = torch.sum(torch.where(used, vols * mean, torch.zeros_like(mean)))
        var_est = torch.sum(torch.where(used, var * vols ** 2 / cnt.clamp(min=1), torch.zeros_like(var)))
        sample = StratifiedSample()
        total_count = vol.total_base_count * multiple
        sample.sub_sizes = [int(c) for c in counts]
        sample.sub_weights = [float(c / v / total_count) for c, v in zip(counts, vols.cpu().numpy())]
        if keep_samples:
            sample.data = config.deliver(xs)
            sample.function_values = config.deliver(values)
        sample.integral = float(int_est)
        sample.integral_err = float(torch.sqrt(var_est))
        return sample
