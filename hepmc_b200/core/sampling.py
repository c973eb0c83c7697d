"""Result containers (mirror of hepmc/core/sampling.py:23-145, host-side and thin).

`data`, `weights`, `function_values` may be NumPy arrays or CUDA tensors; mean / variance
are computed with whichever library holds the data.  The effective sample size and the
bin-wise chi^2 (sampling.py:79-94) run on the device through core/util/stat_tools.py.
"""
import json
import os
import time

import numpy as np
import torch


def _set(*args):
    return all(arg is not None for arg in args)


def _concat(a, b):
    if isinstance(a, torch.Tensor):
        return torch.cat((a, b), dim=0)
    return np.concatenate((a, b), axis=0)


class Sample(object):
    def __init__(self, **kwargs):
        self.data = None
        self.target = None
        self.weights = None
        self._variance = None
        self._mean = None
        self._bin_wise_chi2 = None
        self._effective_sample_size = None
        self._sample_info = [
            ('size', 'data (size)', '%s'),
            ('mean', 'mean', '%s'),
            ('variance', 'variance', '%s'),
            ('bin_wise_chi2', 'bin-wise chi^2', '%.4g, p=%.4g, N=%d'),
            ('effective_sample_size', 'effective sample size', '%s'),
        ]
        for key in kwargs:
            setattr(self, key, kwargs[key])

    def extend_array(self, key, array):
        current = getattr(self, key)
        setattr(self, key, array if current is None else _concat(current, array))

    @property
    def size(self):
        """Number of states per chain: data.shape[0] like the reference; for an ensemble
        (C, T, ndim) the chain length T."""
        try:
            return self.data.shape[-2] if len(self.data.shape) == 3 else self.data.shape[0]
        except AttributeError:
            return None

    @property
    def ndim(self):
        try:
            return self.data.shape[-1] if len(self.data.shape) == 3 else self.data.shape[1]
        except AttributeError:
            return None

    def _states(self):
        """All states as one (N, ndim) array (an ensemble is pooled)."""
        d = self.data
        return d.reshape(-1, d.shape[-1]) if len(d.shape) == 3 else d

    @property
    def mean(self):
        if self._mean is None and _set(self.data):
            d = self._states()
            self._mean = d.mean(dim=0) if isinstance(d, torch.Tensor) else np.mean(d, axis=0)
        return self._mean

    @property
    def variance(self):
        if self._variance is None and _set(self.data):
            d = self._states()
            self._variance = d.var(dim=0, unbiased=False) if isinstance(d, torch.Tensor) else np.var(d, axis=0)
        return self._variance

    @property
    def effective_sample_size(self):
        """Needs the target's known `mean` and `variance` (sampling.py:79-88); None otherwise."""
        if self._effective_sample_size is None and _set(self.data, self.target):
            try:
                moments = self.target.mean, self.target.variance
            except AttributeError:
                return None
            from .util import stat_tools
            self._effective_sample_size = stat_tools.effective_sample_size(self, *moments)
        return self._effective_sample_size

    @property
    def bin_wise_chi2(self):
        if self._bin_wise_chi2 is None and _set(self.data, self.target):
            from .util import stat_tools
            self._bin_wise_chi2 = stat_tools.bin_wise_chi2(self)
        return self._bin_wise_chi2

    def save(self, file_path=None):
        if file_path is None:
            file_path = type(self).__name__ + '-' + str(int(time.time()))
        path, name = os.path.split(file_path)
        info = {entry[0]: repr(getattr(self, entry[0])) for entry in self._sample_info}
        info['type'] = type(self).__name__
        info['target'] = repr(self.target)
        data = self.data.cpu().numpy() if isinstance(self.data, torch.Tensor) else self.data
        np.save(os.path.join(path, name + '-data.npy'), data)
        with open(os.path.join(path, name + '.json'), 'w') as fp:
            json.dump(info, fp, indent=2)

    def _data_table(self):
        titles = [entry[1] for entry in self._sample_info]
        entries = []
        for entry in self._sample_info:
            value = getattr(self, entry[0])
            try:
                entries.append(entry[2] % value)
            except TypeError:
                entries.append('N/A')
        return titles, entries

    def __repr__(self):
        return (type(self).__name__ + '\n\t' + '\n\t'.join(
            '%s: %s' % (t, e) for t, e in zip(*self._data_table())))


class AcceptRejectSampler(object):
    """Acceptance-rejection sampling (mirror of hepmc/core/sampling.py:149-204): proposals
    x ~ sampling_pdf are kept when  u * bound * sampling_pdf(x) <= pdf(x).  Proposals, the
    acceptance test (hepmc_accept_flags) and the compaction run on the device; pdf / sampling /
    sampling_pdf follow the integrand convention (ndim arrays in, one array out; built-in
    densities are evaluated by their kernels, other callables receive CUDA tensors)."""

    def __init__(self, pdf, bound, ndim=1, sampling=None, sampling_pdf=None):
        self.pdf = pdf
        self.c = bound
        self.ndim = ndim
        self.sampling = sampling
        self.sampling_pdf = sampling_pdf

    def sample(self, sample_size, batch=None):
        import ctypes
        from .. import _lib, rng, config
        from .integration import engine
        from .density import as_builtin
        dev = _lib.device()
        sample_size = int(sample_size)
        x = torch.empty((sample_size, self.ndim), dtype=torch.float64, device=dev)
        filled = 0
        dens = as_builtin(self.pdf)
        seed = rng.get_seed()
        while filled < sample_size:
            need = sample_size - filled
            m = int(batch) if batch else max(1024, 2 * need)
            if self.sampling is None:
                prop = engine.uniform_rvs(m, self.ndim)
                q, q_scalar = None, 1.0
            else:
                prop, _ = _lib.to_device(self.sampling(m))
                prop = prop.reshape(m, self.ndim).contiguous()
                qv = self.sampling_pdf(*[prop[:, d] for d in range(self.ndim)])
                if isinstance(qv, (int, float)):
                    q, q_scalar = None, float(qv)
                else:
                    q, q_scalar = _lib.to_device(qv)[0].reshape(-1).contiguous(), 1.0
            if dens is not None:
                f = dens.pdf(prop)
            else:
                f = engine.call_integrand(self.pdf, [prop[:, d] for d in range(self.ndim)])
            flags = torch.empty(m, dtype=torch.uint8, device=dev)
            _lib.call("hepmc_accept_flags", _lib.ptr(f), _lib.ptr(q), q_scalar, float(self.c), m, 0, seed,
                      rng.next_stream(), _lib.ptr(flags), _lib.stream_ptr())
            keep = prop[flags.bool()][:need]
            x[filled:filled + keep.shape[0]] = keep
            filled += keep.shape[0]
        return Sample(data=config.deliver(x), target=self.pdf)
