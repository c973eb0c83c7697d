"""Statistics of a Monte Carlo sample (mirror of hepmc/core/util/stat_tools.py:11-129).

The sums run on the device (csrc/stats.cu): lag autocovariances and the effective sample size
are one CTA per (chain, dimension) with fixed-order reductions; the bin-wise chi^2 is a
histogram kernel (integer atomics), a kernel that throws `int_steps` points into each
populated bin, the target's pdf kernel, and a fixed-order chi^2 reduction.  The host only
prepares the bin edges (np.linspace, as np.histogramdd does) and turns the statistic into a
p-value.  There is no CPU fallback.

Extension over the reference: `data` may be an ensemble (C, T, ndim); the autocovariance
functions then answer per chain, the chi^2 pools all chains.
"""
import ctypes

import numpy as np
import torch

from ... import _lib, rng


def _sample_data(sample):
    return sample.data if hasattr(sample, "data") else sample


def _chains(values):
    """values (T,) | (T, ndim) | (C, T, ndim) -> CUDA tensor (C, T, ndim), came-from-host flag,
    ensemble flag."""
    t, host = _lib.to_device(values)
    if t.dim() == 1:
        t = t[:, None]
    if t.dim() == 2:
        return t[None].contiguous(), host, False
    if t.dim() != 3:
        raise RuntimeError("Bad array shape.")
    return t.contiguous(), host, True


def _moment(value, ndim, dev):
    m = np.broadcast_to(np.asarray(
        value.cpu().numpy() if isinstance(value, torch.Tensor) else value, dtype=np.float64).ravel(), (ndim,))
    return torch.from_numpy(np.array(m, dtype=np.float64)).to(dev)


def auto_cov(values, mean=None, variance=None, max_lag=None):
    """acov[k] = sum_t (x_t - mean)(x_{t+k} - mean) / size for k = 0 .. max_lag-1 (default: all
    `size` lags, like the reference); acov[0] = variance (stat_tools.py:17-37).
    Ensemble data (C, T, ndim) -> (C, max_lag, ndim)."""
    x, host, ens = _chains(values)
    C, T, ndim = x.shape
    if C * ndim > 65535:
        raise RuntimeError("auto_cov: at most 65535 series per call")
    if mean is None or variance is None:
        if ens:
            raise ValueError("auto_cov of an ensemble needs the known mean and variance")
        mean = x[0].mean(dim=0) if mean is None else mean
        variance = x[0].var(dim=0, unbiased=False) if variance is None else variance
    mean_t = _moment(mean, ndim, x.device)
    var_t = _moment(variance, ndim, x.device)
    n_lags = T if max_lag is None else min(int(max_lag), T)
    out = torch.empty((C, n_lags, ndim), dtype=torch.float64, device=x.device)
    _lib.call("hepmc_autocov", _lib.ptr(x), C, T, ndim, _lib.ptr(mean_t), _lib.ptr(var_t), 0, n_lags,
              _lib.ptr(out), _lib.stream_ptr())
    out = out if ens else out[0]
    return _lib.to_host_like(out, host)


def auto_corr(values, mean=None, variance=None, max_lag=None):
    """auto_cov / auto_cov[0] (stat_tools.py:40-43)."""
    acov = auto_cov(values, mean, variance, max_lag)
    return acov / acov[..., :1, :]


def lag_auto_cov(values, k, mean=None):
    """Single lag (stat_tools.py:11-14)."""
    x, host, ens = _chains(values)
    C, T, ndim = x.shape
    if mean is None:
        mean = x.mean(dim=1)[0] if not ens else None
    if mean is None:
        raise ValueError("lag_auto_cov of an ensemble needs the mean")
    mean_t = _moment(mean, ndim, x.device)
    out = torch.empty((C, 1, ndim), dtype=torch.float64, device=x.device)
    _lib.call("hepmc_autocov", _lib.ptr(x), C, T, ndim, _lib.ptr(mean_t), _lib.ptr(mean_t), int(k), 1,
              _lib.ptr(out), _lib.stream_ptr())
    out = out[:, 0] if ens else out[0, 0]
    return _lib.to_host_like(out, host)


def effective_sample_size(sample, mean, var, return_lags=False):
    """size / (1 + 2 sum_lag (1 - lag/size) rho_lag), the sum running from lag 1 while the
    (unbiased) autocorrelation stays >= 0.05 (stat_tools.py:46-70, arXiv:1111.4246).

    :param sample: Sample object or array (T, ndim); ensemble (C, T, ndim) -> (C, ndim).
    :param mean, var: KNOWN moments of the target (do not estimate them from the sample)."""
    x, host, ens = _chains(_sample_data(sample))
    C, T, ndim = x.shape
    mean_t = _moment(mean, ndim, x.device)
    var_t = _moment(var, ndim, x.device)
    ess = torch.empty((C, ndim), dtype=torch.float64, device=x.device)
    lags = torch.empty((C, ndim), dtype=torch.int64, device=x.device)
    _lib.call("hepmc_effective_sample_size", _lib.ptr(x), C, T, ndim, _lib.ptr(mean_t), _lib.ptr(var_t),
              _lib.ptr(ess), _lib.ptr(lags), _lib.stream_ptr())
    if not ens:
        ess, lags = ess[0], lags[0]
    ess, lags = _lib.to_host_like(ess, host), _lib.to_host_like(lags, host)
    return (ess, lags) if return_lags else ess


def _pooled(sample):
    t, host = _lib.to_device(_sample_data(sample))
    if t.dim() == 1:
        t = t[:, None]
    if t.dim() == 3:
        t = t.reshape(-1, t.shape[-1])
    return t.contiguous(), host


def _percentile(sorted_col, q):
    """np.percentile(..., method='linear') of one sorted device column: four scalar reads."""
    n = sorted_col.shape[0]
    pos = q * (n - 1)
    lo = int(np.floor(pos))
    hi = min(lo + 1, n - 1)
    a, b = (float(v) for v in sorted_col[[lo, hi]].cpu())
    t = pos - lo
    diff = b - a
    return b - diff * (1 - t) if t >= 0.5 else a + diff * t


def fd_bins(sample):
    """Freedman-Diaconis bin counts with the size^(-1/(2+d)) scaling (stat_tools.py:73-88)."""
    x, _ = _pooled(sample)
    size, ndim = x.shape
    srt = torch.sort(x, dim=0).values
    mins = srt[0].cpu().numpy()
    maxs = srt[-1].cpu().numpy()
    iqr = np.array([_percentile(srt[:, d], 0.75) - _percentile(srt[:, d], 0.25) for d in range(ndim)])
    widths = 2 * iqr * size ** (-1 / (2 + ndim))
    bins = np.ceil((maxs - mins) / widths).astype(int)
    return np.maximum(1, bins)


def _edges(x, bins, bin_range):
    """Bin edges as np.histogramdd builds them (equal-width from the data range, or from
    `bin_range`; explicit edge arrays pass through)."""
    ndim = x.shape[1]
    try:
        if len(bins) != ndim:
            raise ValueError("The dimension of bins must be equal to the dimension of the sample x.")
    except TypeError:
        bins = ndim * [bins]
    if bin_range is None:
        bin_range = (None,) * ndim
    lo_hi = None
    edges = []
    for d in range(ndim):
        if np.ndim(bins[d]) == 0:
            if bins[d] < 1:
                raise ValueError("`bins[%d]` must be positive, when an integer" % d)
            if bin_range[d] is None:
                if lo_hi is None:
                    lo_hi = (x.min(dim=0).values.cpu().numpy(), x.max(dim=0).values.cpu().numpy()) \
                        if x.shape[0] else (np.zeros(ndim), np.ones(ndim))
                smin, smax = float(lo_hi[0][d]), float(lo_hi[1][d])
            else:
                smin, smax = (float(v) for v in bin_range[d])
            if smin > smax:
                raise ValueError("max must be larger than min in range parameter.")
            if smin == smax:
                smin, smax = smin - 0.5, smax + 0.5
            edges.append(np.linspace(smin, smax, int(bins[d]) + 1))
        else:
            e = np.asarray(bins[d], dtype=np.float64)
            if e.ndim != 1 or e.size < 2 or np.any(e[:-1] > e[1:]):
                raise ValueError("`bins[%d]` must be monotonically increasing, when an array" % d)
            edges.append(e)
    return edges


def histogramdd(sample, bins=10, bin_range=None):
    """np.histogramdd(sample.data, bins, bin_range) on the device -> (counts int64 CUDA tensor of
    shape bins, list of NumPy edge arrays)."""
    x, _ = _pooled(sample)
    edges = _edges(x, bins, bin_range)
    nb = [len(e) - 1 for e in edges]
    total = int(np.prod([int(b) for b in nb], dtype=object))
    if total > 1 << 31:
        raise MemoryError("histogramdd: %d bins do not fit a dense device histogram" % total)
    edges_t = torch.from_numpy(np.concatenate(edges)).to(x.device)
    counts = torch.empty(total, dtype=torch.int64, device=x.device)
    nb_c = (ctypes.c_int * len(nb))(*nb)
    _lib.call("hepmc_histogramdd", _lib.ptr(x), x.shape[0], x.shape[1], _lib.ptr(edges_t), nb_c,
              _lib.ptr(counts), _lib.stream_ptr())
    return counts.reshape(nb), edges


def _target_pdf(sample):
    target = sample.target
    if hasattr(target, "pdf"):
        return target.pdf
    return lambda xs: target(*xs.T)


def bin_wise_chi2(sample, bins=None, bin_range=None, min_count=10, int_steps=100, replay=None):
    """Bin-wise chi^2 / dof of the sample against its target (stat_tools.py:91-129).

    Bins with at least `min_count` entries are compared with the expected count
    size * vol * <pdf>_bin, the bin average taken over `int_steps` uniform points; bins whose
    expectation is below `min_count` are dropped as well.  Returns (chi2 / (n - 1), p, n) or
    (None, None, None) when no bin qualifies.

    :param replay: uniforms (n_populated_bins, int_steps, ndim) consumed instead of the native
        generator (the reference's np.random.uniform draws, in bin order)."""
    x, _ = _pooled(sample)
    size, ndim = x.shape
    if bins is None:
        bins = fd_bins(x)
    counts, edges = histogramdd(x, bins, bin_range)
    nb = [len(e) - 1 for e in edges]
    flat_counts = counts.reshape(-1)
    flat = torch.nonzero(flat_counts >= min_count).reshape(-1).contiguous()   # ascending = np.where order
    m = int(flat.shape[0])
    out = torch.zeros(2, dtype=torch.float64, device=x.device)
    if m > 0:
        edges_t = torch.from_numpy(np.concatenate(edges)).to(x.device)
        nb_c = (ctypes.c_int * len(nb))(*nb)
        u = None
        if replay is not None:
            u, _ = _lib.to_device(replay)
            if tuple(u.shape) != (m, int_steps, ndim):
                raise ValueError("replay must have shape %s" % ((m, int_steps, ndim),))
        pts = torch.empty((m * int_steps, ndim), dtype=torch.float64, device=x.device)
        _lib.call("hepmc_bin_points", ndim, _lib.ptr(edges_t), nb_c, _lib.ptr(flat), m, int(int_steps),
                  rng.get_seed(), rng.next_stream(), _lib.ptr(u), _lib.ptr(pts), _lib.stream_ptr())
        pdf = _target_pdf(sample)(pts)
        pdf, _ = _lib.to_device(pdf)
        pdf = pdf.reshape(-1).contiguous()
        if pdf.shape[0] != m * int_steps:
            raise RuntimeError("target pdf returned %d values for %d points" % (pdf.shape[0], m * int_steps))
        vol = float(np.prod([e[1] - e[0] for e in edges]))
        expected = torch.empty(m, dtype=torch.float64, device=x.device)
        _lib.call("hepmc_bin_chi2", _lib.ptr(pdf), _lib.ptr(flat), _lib.ptr(flat_counts), m, int(int_steps),
                  float(size), vol, float(min_count), _lib.ptr(expected), _lib.ptr(out), _lib.stream_ptr())
    chi2, n = (float(v) for v in out.cpu())
    n = int(n)
    if n == 0:
        return None, None, None
    dof = n - 1
    if dof == 0:
        return float('nan') if chi2 == 0 else float('inf'), float('nan'), n
    p = float(torch.special.gammaincc(torch.tensor(dof / 2.0, dtype=torch.float64),
                                      torch.tensor(chi2 / 2.0, dtype=torch.float64)))
    return chi2 / dof, p, n
