"""Separable grid on the unit hypercube: the VEGAS sampling distribution
(hepmc/core/integration/stratified_volume.py:11-182), and the partition of StratifiedMC
(per-bin base counts, `iterate`).

State = per axis the bin boundaries `bounds[d]` (K+1) and widths `sizes[d]` (K).  A bin is
chosen uniformly per axis and the point uniformly inside it; pdf = 1 / (K^ndim * volume).
The state lives in ONE device tensor `(ndim, 2K+1)` = [bounds | sizes] that the kernels
read and `hepmc_vegas_update_grid` rewrites in place; the `bounds` / `sizes` properties
expose it as lists of NumPy arrays like the reference.
"""
import ctypes

import numpy as np
import torch

from ..density import Distribution
from ..util import interpret_array
from ... import _lib, rng, config
from . import engine


class GridVolumes(Distribution):

    def __init__(self, bounds=None, base_counts=None, default_base_count=1, ndim=1, divisions=1):
        # {multi-index tuple: base number of samples}; bins not listed use default_base_count (:13-37)
        self.base_counts = dict(base_counts) if base_counts else dict()
        self.default_base_count = default_base_count
        self._grid = None       # device tensor (ndim, 2K+1), created on first device use
        self._host = None       # float64 array (ndim, 2K+1): the authoritative copy until then
        self._ragged = None     # list of per-axis bounds when the axes have different divisions (host only)
        if bounds is None:
            Distribution.__init__(self, ndim, False)
            if not np.isscalar(divisions):
                if len(divisions) != ndim:
                    raise RuntimeError("Divisions must be scalar or of length ndim")
                self.bounds = [np.linspace(0, 1, int(k) + 1) for k in divisions]
            else:
                self.bounds = [np.linspace(0, 1, divisions + 1) for _ in range(ndim)]
        else:
            Distribution.__init__(self, len(bounds), False)
            self.bounds = bounds
        self.initial_bounds = [np.copy(b) for b in self.bounds]

    # ---- state ------------------------------------------------------------------
    @property
    def divisions(self):
        """Bins per axis (an int for the regular grids VEGAS uses, a tuple when the axes differ)."""
        if self._ragged is not None:
            return tuple(len(b) - 1 for b in self._ragged)
        return (self._state_host().shape[1] - 1) // 2

    def _shape(self):
        k = self.divisions
        return tuple(k) if isinstance(k, tuple) else (k,) * self.ndim

    def _state_host(self):
        if self._ragged is not None:
            raise NotImplementedError("the grid kernels (VEGAS, hepmc_grid_pdf) need the same number of divisions on "
                                      "every axis; grids with different divisions serve StratifiedMC / pdf / rvs")
        if self._grid is not None:
            self._host = self._grid.cpu().numpy()
        return self._host

    def _set_state(self, host):
        self._ragged = None
        self._host = np.ascontiguousarray(host, dtype=np.float64)
        self._grid = None
        k = (self._host.shape[1] - 1) // 2
        self.partition_count = float(k) ** self.ndim   # floating point: no int64 overflow (App. B1)
        # _update_total_base_count, :149-155
        self.total_base_count = (sum(self.base_counts.values()) +
                                 self.default_base_count * (self.partition_count - len(self.base_counts)))

    def device_grid(self):
        """The (ndim, 2K+1) CUDA tensor the kernels use (authoritative once created)."""
        if self._ragged is not None:
            self._state_host()      # raises: the grid kernels need equal divisions on every axis
        if self._grid is None:
            self._grid, _ = _lib.to_device(self._host)
        return self._grid

    def reset(self):
        self.bounds = [np.copy(b) for b in self.initial_bounds]

    @property
    def bounds(self):
        if self._ragged is not None:
            return [b.copy() for b in self._ragged]
        st = self._state_host()
        k = self.divisions
        return [st[d, :k + 1].copy() for d in range(self.ndim)]

    @bounds.setter
    def bounds(self, bounds):
        rows = [np.asarray(x, dtype=np.float64) for x in bounds]
        if len(set(len(r) for r in rows)) > 1:
            self._ragged = [r.copy() for r in rows]
            self._host = self._grid = None
            self.partition_count = float(np.prod([len(r) - 1 for r in rows]))
            self.total_base_count = (sum(self.base_counts.values()) +
                                     self.default_base_count * (self.partition_count - len(self.base_counts)))
            return
        b = np.stack(rows)
        self._set_state(np.concatenate([b, np.diff(b, axis=1)], axis=1))

    @property
    def sizes(self):
        if self._ragged is not None:
            return [np.diff(b) for b in self._ragged]
        st = self._state_host()
        k = self.divisions
        return [st[d, k + 1:].copy() for d in range(self.ndim)]

    @sizes.setter
    def sizes(self, sizes):
        rows = [np.asarray(x, dtype=np.float64) for x in sizes]
        self.bounds = [np.concatenate([[0.0], np.cumsum(r)]) for r in rows]   # bounds[d][0] = 0 (App. B2)

    # ---- sampling ---------------------------------------------------------------
    def _sample(self, count, want_idx):
        dev = _lib.device()
        k = self.divisions
        x = torch.empty((self.ndim, count), dtype=torch.float64, device=dev)
        w = torch.empty(count, dtype=torch.float64, device=dev)
        idx = torch.empty((self.ndim, count), dtype=torch.int64, device=dev)
        if count:
            blk = engine.none_block(self.ndim)
            _lib.call("hepmc_vegas_sample", ctypes.byref(blk), _lib.ptr(self.device_grid()), self.ndim, k,
                      count, 0, _lib.default_chunk(count), rng.get_seed(), rng.next_stream(), rng.mode_code(),
                      None, None, _lib.ptr(x), None, _lib.ptr(w), _lib.ptr(idx), None, None, _lib.stream_ptr())
        return x, w, idx

    def _bins_dev(self):
        dev = _lib.device()
        return [torch.as_tensor(b, device=dev) for b in self.bounds]

    def random_bins(self, count):
        """(ndim, count) uniformly chosen bin indices (stratified_volume.py:104-113)."""
        if self._ragged is not None:
            u = engine.uniform_rvs(int(count), self.ndim)
            ks = torch.as_tensor(self._shape(), device=u.device, dtype=torch.float64)
            return config.deliver(torch.minimum((u * ks).floor(), ks - 1).to(torch.int64).t().contiguous())
        return config.deliver(self._sample(int(count), True)[2])

    def sample_indices(self, indices):
        """Points uniform inside the given bins, TRANSPOSED (ndim, N) like the reference
        (:115-122): replayed through the sampling kernel with fresh native uniforms."""
        idx, host = _lib.to_device(indices, dtype=torch.int64)
        n = idx.shape[1]
        u = engine.uniform_rvs(n * self.ndim, 1).reshape(self.ndim, n).contiguous()
        if self._ragged is not None:
            b = self._bins_dev()
            x = torch.stack([b[d][idx[d]] + (b[d][idx[d] + 1] - b[d][idx[d]]) * u[d] for d in range(self.ndim)])
            return _lib.to_host_like(x, host)
        x = torch.empty((self.ndim, n), dtype=torch.float64, device=idx.device)
        w = torch.empty(n, dtype=torch.float64, device=idx.device)
        io = torch.empty((self.ndim, n), dtype=torch.int64, device=idx.device)
        if n:
            blk = engine.none_block(self.ndim)
            _lib.call("hepmc_vegas_sample", ctypes.byref(blk), _lib.ptr(self.device_grid()), self.ndim,
                      self.divisions, n, 0, _lib.default_chunk(n), 0, 0, 0, _lib.ptr(idx), _lib.ptr(u),
                      _lib.ptr(x), None, _lib.ptr(w), _lib.ptr(io), None, None, _lib.stream_ptr())
        return _lib.to_host_like(x, host)

    def rvs(self, sample_count):
        """(N, ndim) points distributed like the grid (:124-126)."""
        if self._ragged is not None:
            idx = self.random_bins(int(sample_count))
            x = self.sample_indices(idx)
            x = x if isinstance(x, torch.Tensor) else torch.as_tensor(x, device=_lib.device())
            return config.deliver(x.t().contiguous())
        x, _, _ = self._sample(int(sample_count), False)
        return config.deliver(x.t().contiguous())

    # ---- stratified partition (StratifiedMC) ---------------------------------------
    def get_count(self, index, multiple=1):
        """Samples of bin `index` (:127-134; a non-integer multiple truncates)."""
        return int(multiple * self.base_counts.get(tuple(int(i) for i in index), self.default_base_count))

    def _cells(self, multiple=1):
        """All bins in np.ndindex order: counts (n_cells,), lower / upper corners (n_cells, ndim)."""
        shape = self._shape()
        b = self.bounds
        index = np.stack(np.unravel_index(np.arange(int(np.prod(shape))), shape), axis=1)
        counts = np.full(index.shape[0], int(multiple * self.default_base_count), dtype=np.int64)
        for key, value in self.base_counts.items():
            counts[np.ravel_multi_index(tuple(int(i) for i in key), shape)] = int(multiple * value)
        lower = np.stack([b[d][index[:, d]] for d in range(self.ndim)], axis=1)
        upper = np.stack([b[d][index[:, d] + 1] for d in range(self.ndim)], axis=1)
        return counts, lower, upper

    def _cell_points(self, counts, lower, upper, replay=None):
        """One batch for all bins: uniforms from the native generator (or the replay tape), mapped
        into each bin; returns the (N, ndim) CUDA tensor and the bin id of every point."""
        dev = _lib.device()
        n = int(counts.sum())
        if replay is not None:
            u, _ = _lib.to_device(np.concatenate([np.asarray(r, dtype=np.float64).reshape(-1, self.ndim) for r in replay]))
            if u.shape[0] != n:
                raise ValueError("replay holds %d points, the partition needs %d" % (u.shape[0], n))
        else:
            u = engine.uniform_rvs(n, self.ndim)
        cell = torch.repeat_interleave(torch.arange(counts.shape[0], device=dev), torch.as_tensor(counts, device=dev))
        lo = torch.as_tensor(lower, device=dev)[cell]
        hi = torch.as_tensor(upper, device=dev)[cell]
        return lo + (hi - lo) * u, cell                      # :143 lower + (upper - lower) * rand

    def iterate(self, multiple=1, replay=None):
        """Yield (count, samples (count, ndim), volume) bin by bin (:136-147)."""
        counts, lower, upper = self._cells(multiple)
        xs, _ = self._cell_points(counts, lower, upper, replay)
        start = 0
        for c, lo, hi in zip(counts, lower, upper):
            yield int(c), config.deliver(xs[start:start + int(c)]), float(np.prod(hi - lo))
            start += int(c)

    # ---- density ----------------------------------------------------------------
    def _counts(self, idx):
        """Base count of the bins with (ndim, N) indices, on any device (stratified_volume.py:57-65):
        default_base_count, overwritten for the listed bins -- a default of 0 with explicit counts is valid."""
        cnt = torch.full((idx.shape[1],), float(self.default_base_count), dtype=torch.float64, device=idx.device)
        for key, value in self.base_counts.items():
            hit = torch.ones(idx.shape[1], dtype=torch.bool, device=idx.device)
            for d in range(self.ndim):
                hit &= idx[d] == int(key[d])
            cnt = torch.where(hit, torch.full_like(cnt, float(value)), cnt)
        return cnt

    def pdf_indices(self, indices):
        """count / total_base_count / volume of the given bins (:67-72)."""
        idx = torch.as_tensor(np.asarray(indices) if not isinstance(indices, torch.Tensor) else indices).cpu().long()
        sz = [torch.as_tensor(v) for v in self.sizes]
        vols = torch.ones(idx.shape[1], dtype=torch.float64)
        for d in range(self.ndim):
            vols = vols * sz[d][idx[d]]
        out = self._counts(idx) / self.total_base_count / vols
        return out.numpy()

    def pdf(self, xs):
        """1 / (K^ndim * bin volume) inside the open unit cube, 0 outside (:74-83)."""
        xs = interpret_array(xs, self.ndim)
        x_dev, host = _lib.to_device(xs)
        if self._ragged is not None:
            # different divisions per axis: bin search with torch ops on the device
            b = self._bins_dev()
            shape = self._shape()
            idx = torch.stack([(torch.bucketize(x_dev[:, d].contiguous(), b[d], right=True) - 1).clamp_(0, shape[d] - 1)
                               for d in range(self.ndim)])
            vols = torch.ones(x_dev.shape[0], dtype=torch.float64, device=x_dev.device)
            for d in range(self.ndim):
                vols = vols * (b[d][idx[d] + 1] - b[d][idx[d]])
            inside = ((x_dev > 0) & (x_dev < 1)).all(dim=1)
            out = self._counts(idx) / self.total_base_count / vols
            return _lib.to_host_like(torch.where(inside, out, torch.zeros_like(out)), host)
        out = torch.empty(x_dev.shape[0], dtype=torch.float64, device=x_dev.device)
        if x_dev.shape[0]:
            _lib.call("hepmc_grid_pdf", _lib.ptr(self.device_grid()), self.ndim, self.divisions, _lib.ptr(x_dev),
                      x_dev.shape[0], _lib.ptr(out), _lib.stream_ptr())
            if self.base_counts:
                # the kernel answers 1 / (K^ndim vol); bins with their own base count are rescaled
                grid = self.device_grid()
                k = self.divisions
                idx = torch.stack([torch.bucketize(x_dev[:, d].contiguous(), grid[d, :k + 1].contiguous(), right=True) - 1
                                   for d in range(self.ndim)]).clamp_(0, k - 1)
                out = out * (self.partition_count / self.total_base_count) * self._counts(idx)
        return _lib.to_host_like(out, host)

    def pdf_gradient(self, xs):
        raise NotImplementedError("Density is a step function.")
