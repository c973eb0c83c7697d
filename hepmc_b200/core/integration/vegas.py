"""VEGAS importance-sampling integration (hepmc/core/integration/vegas.py:14-144).

Per iteration j:  uniform bin + offset per axis -> point, weight w = prod(sizes) * K^ndim,
wf = f * w;  est_j = mean(wf), var_j = var0(wf)/N;  per-(axis, bin) first moments drive
the damped grid update  inv = K*est[d,k] * c/(j+1+c) + (est_j/size) * (j+1)/(j+1+c).
The whole iteration is four launches (fused sample kernel, two-level ordered slot reduction into
packed slot rows, grid update) with no host synchronisation; under torch.distributed one
all-gather of the 8 packed slot rows ((3 + ndim*K) doubles each) is the only exchange.
"""
import ctypes
import math

import numpy as np
import torch

from .integration import IntegrationSample
from .stratified_volume import GridVolumes
from ..density import as_builtin
from ... import _lib, rng, parallel, config
from . import engine


class VegasSample(IntegrationSample):

    def __init__(self, **kwargs):
        self.chi2 = None
        self.estimates = None    # est_j per iteration
        self.variances = None    # var_j per iteration
        super().__init__(**kwargs)


class VegasMC(object):

    def __init__(self, ndim=1, divisions=1, c=3, name="MC VEGAS", var_weighted=False):
        if not (1 <= int(ndim) <= _lib.MAX_DIM):
            raise ValueError("VegasMC supports 1 <= ndim <= %d (got %r)" % (_lib.MAX_DIM, ndim))
        if not (1 <= int(divisions) <= 255):
            raise ValueError("VegasMC supports 1 <= divisions <= 255 (bin ids are bytes in the kernels; got %r)" % (divisions,))
        self.ndim = ndim
        self.volumes = GridVolumes(ndim=ndim, divisions=divisions)
        self.divisions = divisions
        self.c = c
        self.var_weighted = var_weighted
        self.method_name = name

    def get_interface_infer_multiple(self, sub_eval_count):
        """Interface taking a total evaluation budget (vegas.py:40-66)."""
        def interface(fn, eval_count, apriori=True, chi=False):
            iterations = eval_count // sub_eval_count
            return self(fn, sub_eval_count, iterations, apriori, chi)
        interface.method_name = self.method_name
        return interface

    def weights_at(self, indices):
        """prod_d sizes[d][indices[d]] * K^ndim (vegas.py:68-72); small host helper."""
        sizes = self.volumes.sizes
        idx = indices.cpu().numpy() if isinstance(indices, torch.Tensor) else np.asarray(indices)
        vols = np.prod([sizes[d][idx[d]] for d in range(self.ndim)], axis=0)
        return vols * self.volumes.partition_count

    def __call__(self, fn, sub_eval_count, iterations, apriori=True, chi=False,
                 keep_samples=True, replay=None):
        """:param fn: integrand (a built-in density is fused into the sampling kernel).
        :param sub_eval_count: points per iteration (whole job, all GPUs).
        :param iterations: number of adaptation iterations.
        :param apriori: reset the grid first.
        :param chi: also compute chi^2/dof of the per-iteration estimates.
        :param keep_samples: keep data (ndim*iterations, N) [the reference's transposed
            layout, vegas.py:105], function_values and weights; switch off for large N.
        :param replay: (idx, u) tapes of shape (iterations, ndim, N): the reference's own
            randint / rand draws, consumed instead of the native generator."""
        _lib.require_cuda()
        if apriori:
            self.volumes.reset()
        assert not chi or iterations > 1, "Can only compute chi^2 if there is more than one iteration "
        n_total = int(sub_eval_count)
        ndim, K = self.ndim, int(self.volumes.divisions)
        dev = _lib.device()
        rank, ws = parallel.world()
        dens = as_builtin(fn)
        fused = dens is not None
        if fused and dens.ndim != ndim:
            raise RuntimeWarning("Mismatching dimensions.")
        idx_tape = u_tape = None
        if replay is not None:
            if ws != 1:
                raise RuntimeError("replay mode is single-process")
            idx_tape, _ = _lib.to_device(replay[0], dtype=torch.int64)
            u_tape, _ = _lib.to_device(replay[1])
            if tuple(idx_tape.shape) != (iterations, ndim, n_total) or idx_tape.shape != u_tape.shape:
                raise ValueError("replay tapes must have shape (iterations, ndim, sub_eval_count)")
        grid = self.volumes.device_grid()
        est_var = torch.zeros((iterations, 2), dtype=torch.float64, device=dev)
        if fused and ws == 1 and replay is None and not keep_samples and iterations > 0:
            # whole call in ONE library entry (hepmc_vegas_integrate): all iterations enqueued back
            # to back, no Python between launches -- matters at small N (launch-latency bound)
            blk = dens._block()
            nbytes = int(_lib.load().hepmc_vegas_workspace_bytes(ndim, K, n_total))
            work = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device=dev)
            _lib.call("hepmc_vegas_integrate", ctypes.byref(blk), _lib.ptr(grid), ndim, K, n_total, int(iterations),
                      float(self.c), rng.get_seed(), rng.next_stream(iterations), rng.mode_code(), _lib.ptr(est_var), _lib.ptr(work),
                      nbytes, _lib.stream_ptr())
            return self._combine(est_var, n_total, iterations, chi, None)
        chunk = _lib.default_chunk(n_total)
        first, n_local, begins = parallel.local_span(n_total, chunk)
        nc_local, n_slots_local = begins[-1], len(begins) - 1
        cols = ndim * K
        stats_p = torch.zeros((max(nc_local, 1), 3), dtype=torch.float64, device=dev)
        hist_p = torch.zeros((max(nc_local, 1), cols), dtype=torch.float64, device=dev)
        # packed slot rows [count, mean, M2, hist...]: written by the fused slot reduction, all-gathered as
        # they are, read in place by the grid update (no slicing or copying between the three)
        slot_local = torch.zeros((n_slots_local, 3 + cols), dtype=torch.float64, device=dev)
        slots_all = slot_local if ws == 1 else torch.empty((_lib.N_SLOTS, 3 + cols), dtype=torch.float64, device=dev)
        scratch = torch.zeros(int(_lib.load().hepmc_reduce_slots_scratch_bytes(n_slots_local, cols)) // 8 + 1,
                              dtype=torch.float64, device=dev)    # zero: arrival counters of the fused reduction
        begins_c = _lib.i64_array(begins)
        # multi-GPU: the slot rows travel over NVLink peer memory inside the reduction / update kernels when a
        # symmetric allocation is available (parallel.SlotExchange), else through one NCCL all-gather
        exchange = parallel.SlotExchange.get(cols) if (ws > 1 and dev.type == "cuda") else None
        slot0 = parallel.slot_span()[0] if ws > 1 else 0
        seed = rng.get_seed()
        sid0 = rng.next_stream(iterations)
        blk = dens._block() if fused else engine.none_block(ndim)
        need_pts = keep_samples or not fused
        kept = None
        if keep_samples and iterations:
            # the kernels write every iteration straight into its slice of the returned arrays
            # (data is (ndim * iterations, N): the reference's transposed layout, vegas.py:105)
            kept = {"data": torch.empty((iterations * ndim, n_local), dtype=torch.float64, device=dev),
                    "function_values": torch.empty(iterations * n_local, dtype=torch.float64, device=dev),
                    "weights": torch.empty(iterations * n_local, dtype=torch.float64, device=dev)}
        st = _lib.stream_ptr()
        for j in range(iterations):
            if kept is not None:
                x = kept["data"][j * ndim:(j + 1) * ndim]
                w = kept["weights"][j * n_local:(j + 1) * n_local]
                f = kept["function_values"][j * n_local:(j + 1) * n_local] if fused else None
            else:
                x = torch.empty((ndim, n_local), dtype=torch.float64, device=dev) if need_pts else None
                w = torch.empty(n_local, dtype=torch.float64, device=dev) if need_pts else None
                f = None
            idx = torch.empty((ndim, n_local), dtype=torch.int64, device=dev) if not fused else None
            if n_local:
                _lib.call("hepmc_vegas_sample", ctypes.byref(blk), _lib.ptr(grid), ndim, K, n_local, first, chunk,
                          seed, sid0 + j, rng.mode_code(),
                          _lib.ptr(idx_tape[j]) if idx_tape is not None else None,
                          _lib.ptr(u_tape[j]) if u_tape is not None else None,
                          _lib.ptr(x), _lib.ptr(f), _lib.ptr(w), _lib.ptr(idx),
                          _lib.ptr(stats_p), _lib.ptr(hist_p), st)
                if not fused:
                    f = engine.call_integrand(fn, [x[d] for d in range(ndim)])
                    _lib.call("hepmc_vegas_accumulate", _lib.ptr(f), _lib.ptr(w), _lib.ptr(idx), ndim, K,
                              n_local, chunk, _lib.ptr(stats_p), _lib.ptr(hist_p), st)
            if exchange is not None:
                epoch = exchange.next_epoch()
                _lib.call("hepmc_reduce_slots_p2p", _lib.ptr(stats_p), _lib.ptr(hist_p), begins_c, n_slots_local, cols,
                          _lib.ptr(slot_local), _lib.ptr(scratch), exchange.peer_ptrs, ws, slot0, epoch, st)
                _lib.call("hepmc_vegas_update_grid_p2p", _lib.ptr(exchange.buffer), cols, epoch, _lib.ptr(grid), ndim, K,
                          float(self.c), j, _lib.ptr(est_var), _lib.ptr(exchange.timeout), st)
            else:
                _lib.call("hepmc_reduce_slots", _lib.ptr(stats_p), _lib.ptr(hist_p), begins_c, n_slots_local, cols,
                          _lib.ptr(slot_local), _lib.ptr(scratch), st)
                if ws > 1:
                    parallel.gather_slots(slot_local, out=slots_all)
                _lib.call("hepmc_vegas_update_grid_packed", _lib.ptr(slots_all), slots_all.shape[0],
                          _lib.ptr(grid), ndim, K, float(self.c), j, _lib.ptr(est_var), st)
            if kept is not None and not fused and n_local:
                kept["function_values"][j * n_local:(j + 1) * n_local] = f
        out = self._combine(est_var, n_total, iterations, chi, kept)
        if exchange is not None:
            exchange.check()
        return out

    def _combine(self, est_var, n_total, iterations, chi, kept):
        """Combination of the iterations, vegas.py:124-142 (host scalars)."""
        sample = VegasSample()
        ev = est_var.cpu().numpy()            # the only synchronisation of the call
        est_j, var_j = ev[:, 0], ev[:, 1]
        n = float(n_total)
        if self.var_weighted:
            norm = np.sum(n / var_j)
            total_est = np.sum(n * est_j / var_j) / norm
            var = np.sum(n ** 2 / var_j) / norm ** 2
        else:
            total_est = np.sum(n * est_j) / (iterations * n)
            var = np.sum(n ** 2 * var_j) / (iterations * n) ** 2
        sample.integral = float(total_est)
        sample.integral_err = float(np.sqrt(var))
        sample.estimates, sample.variances = est_j, var_j
        if chi:
            sample.chi2 = float(np.sum((est_j - total_est) ** 2 / var_j) / (iterations - 1))
        if kept is not None and iterations:
            sample.data = config.deliver(kept["data"])
            sample.function_values = config.deliver(kept["function_values"])
            sample.weights = config.deliver(kept["weights"])
        return sample
