"""Binned KS statistic of the reference's test harness (test infrastructure).

Restates /root/reference/testsuit/testing/stats.py:36-53 (`ks`): per axis, the
max distance between the cumulative marginals of two histograms, and the
critical value d_alpha = sqrt(ln(2/alpha) / (2 n_a)).
"""
import numpy as np


def ks(a, b, alpha=0.05):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    ndim = a.ndim
    axes = np.arange(ndim)
    out = []
    for axis in range(ndim):
        n_a = np.sum(a)
        n_b = np.sum(b)
        ps = np.sum(a, axis=tuple(axes[axes != axis]))
        pt = np.sum(b, axis=tuple(axes[axes != axis]))
        cps = np.cumsum(ps) / n_a
        cpt = np.cumsum(pt) / n_b
        out.append([np.max(np.abs(cps - cpt)), np.sqrt(np.log(2 / alpha) / (2 * n_a))])
    return out
