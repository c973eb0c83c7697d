"""Importance-sampling Monte Carlo integration (hepmc/core/integration/importance.py:5-51)
and multi-channel Monte Carlo (importance.py:54-230).

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


class MultiChannelMC(object):
    """Multi-channel Monte Carlo integration with adaptive channel weights
    (importance.py:54-230).  Phase 1: adapt, discard estimates; phase 2: adapt and estimate;
    phase 3: estimate with frozen weights.  One iteration = one fused kernel (channel pick,
    draw, mixture density, integrand, f/p moments, per-channel second moments) + the ordered
    slot reduction; the weight update  alpha_c <- alpha_c W_c^b / sum  runs on the host scalars.
    """

    def __init__(self, channels, b=.5, name="MC Multi C.", var_weighted=False):
        self.method_name = name
        self.channels = channels
        self.var_weighted = var_weighted
        self.b = b

    def get_interface_ratios(self, sub_eval_count=100, r1=0, r2=1, r3=0):
        if r1 < 0 or r2 < 0 or r3 < 0:
            raise ValueError("Ratios cannot be smaller than 0: %d, %d, %d." % (r1, r2, r3))
        if not np.isclose(1, r1 + r2 + r3):
            raise ValueError("Ratios must sum to 1: %d + %d + %d = %d." % (r1, r2, r3, r1 + r2 + r3))

        def interface(fn, eval_count, apriori=True):
            num_iterations = eval_count // sub_eval_count
            m1, m2, m3 = int(r1 * num_iterations), int(r2 * num_iterations), int(r3 * num_iterations)
            remaining = eval_count - (m1 + m2 + m3) * sub_eval_count
            iter_3 = [sub_eval_count] * m3
            if remaining:
                iter_3.append(remaining)
            return self(fn, [sub_eval_count] * m1, [sub_eval_count] * m2, iter_3, apriori=apriori)
        interface.method_name = self.method_name
        return interface

    def iterate(self, fn, eval_count, update_weights=True, get_estimate=True, keep_samples=True, replay=None):
        """One iteration (importance.py:141-175).  Returns (data, fn_values, weights, estimate,
        w_est) when get_estimate, like the reference (arrays are None with keep_samples=False).
        replay = (xs (N, ndim), channel_of_origin (N,)) feeds the reference's own sample."""
        import ctypes
        from . import engine as eng
        from ... import parallel
        mc = self.channels
        n_total = int(eval_count)
        dens = as_builtin(fn)
        fused = dens is not None
        blk = dens._block() if fused else eng.none_block(mc.ndim)
        alpha = mc.channels_weight.copy()
        x, f, p, ch, stats_p, ch_p, begins, chunk = mc._launch(
            blk, n_total, (keep_samples, keep_samples, keep_samples, keep_samples), replay)
        if not fused:
            f = eng.call_integrand(fn, [x[:, d] for d in range(mc.ndim)])
            if x.shape[0]:
                _lib.call("hepmc_multichannel_accumulate", _lib.ptr(f), _lib.ptr(p), _lib.ptr(ch), mc.count,
                          x.shape[0], chunk, _lib.ptr(stats_p), _lib.ptr(ch_p), _lib.stream_ptr())
        # ordered slot reduction (+ all-gather under torch.distributed), then a small D2H copy
        n_slots = len(begins) - 1
        dev = stats_p.device
        slot = torch.zeros((n_slots, 3 + 2 * mc.count), dtype=torch.float64, device=dev)
        s_stats = torch.zeros((n_slots, 3), dtype=torch.float64, device=dev)
        s_ch = torch.zeros((n_slots, 2 * mc.count), dtype=torch.float64, device=dev)
        _lib.call("hepmc_reduce_stats", _lib.ptr(stats_p), _lib.i64_array(begins), n_slots, _lib.ptr(s_stats), _lib.stream_ptr())
        _lib.call("hepmc_reduce_hist", _lib.ptr(ch_p), _lib.i64_array(begins), n_slots, 2 * mc.count, _lib.ptr(s_ch), _lib.stream_ptr())
        slot[:, :3] = s_stats
        slot[:, 3:] = s_ch
        slots = parallel.gather_slots(slot)
        total = torch.empty(3, dtype=torch.float64, device=dev)
        all_stats = slots[:, :3].contiguous()
        _lib.call("hepmc_reduce_stats", _lib.ptr(all_stats), _lib.i64_array([0, slots.shape[0]]), 1, _lib.ptr(total), _lib.stream_ptr())
        per_ch = slots[:, 3:].sum(dim=0).cpu().numpy()      # 8 rows: order-fixed, tiny
        n, mean, m2 = total.cpu().tolist()
        counts, sumsq = per_ch[:mc.count], per_ch[mc.count:]
        active = np.where(counts > 0)[0]
        w_fn = sumsq[active] / counts[active]                 # importance.py:166-167
        a_act = alpha[active]
        if update_weights:
            factors = a_act * np.power(w_fn, self.b)           # :170-172
            mc.update_channel_weights(factors / np.sum(factors), active)
        if get_estimate:
            estimate = mean                                    # np.mean(weighted)
            w_est = float(np.sum(a_act * w_fn))
            if keep_samples:
                xs, fs, ps, sizes = mc._grouped(x, ch, f, p)
                return config.deliver(xs), config.deliver(fs), config.deliver(ps), estimate, w_est
            return None, None, None, estimate, w_est

    def __call__(self, fn, eval_count_1, eval_count_2, eval_count_3, apriori=True, keep_samples=True, replay=None):
        """importance.py:177-230.  replay: list of (xs, channel ids), one per iteration of all phases."""
        _lib.require_cuda()
        if apriori:
            self.channels.reset()
        it = 0
        for eval_count in eval_count_1:
            self.iterate(fn, eval_count, True, False, False, replay[it] if replay else None)
            it += 1
        sample = IntegrationSample()
        eval_counts = np.concatenate([eval_count_2, eval_count_3]).astype(int)
        m2 = len(eval_count_2)
        ws_est = np.empty(eval_counts.size)
        estimates = np.empty(eval_counts.size)
        kept = []
        for j, eval_count in enumerate(eval_counts):
            xs, values, weights, est, w = self.iterate(fn, eval_count, j < m2, True, keep_samples,
                                                       replay[it] if replay else None)
            it += 1
            estimates[j], ws_est[j] = est, w
            if keep_samples:
                kept.append((values, weights, xs))
        if kept:
            # one concatenation at the end (extend_array per iteration copies everything again each time)
            cat = (lambda parts: torch.cat(parts, dim=0)) if isinstance(kept[0][2], torch.Tensor) else \
                (lambda parts: np.concatenate(parts, axis=0))
            sample.function_values, sample.weights, sample.data = (cat([k[i] for k in kept]) for i in range(3))
        if eval_counts.size:
            if self.var_weighted:
                variances = (ws_est - estimates ** 2) / eval_counts
                norm = np.sum(eval_counts / variances)
                total_est = np.sum(eval_counts * estimates / variances) / norm
                var = np.sum(eval_counts ** 2 / variances) / norm ** 2
            else:
                total_evaluations = np.sum(eval_counts)
                total_est = np.sum(estimates * eval_counts) / total_evaluations
                var = (np.sum(eval_counts * ws_est / total_evaluations) - total_est ** 2) / total_evaluations
            sample.integral = float(total_est)
            sample.integral_err = float(np.sqrt(var))
        return sample
