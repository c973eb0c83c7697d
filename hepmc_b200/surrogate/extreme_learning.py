"""Extreme-learning-machine surrogate (mirror of hepmc/surrogate/extreme_learning.py).

A random hidden layer of `node_count` basis functions with least-squares output weights:
z(x) = bias + sum_i ow_i g_i(x).  `GaussianBasis`: g_i = exp(-|x-c_i|^2 / (2 w_i^2));
`TrigBasis`: g_i = cos(b_i + W_i . x).  `eval`, `eval_gradient` and `output_matrix` run in
hepmc_elm_eval; inside HMC the gradient is evaluated in the chain kernel from shared-memory
node tables (precision f64 / f32) or on the tcgen05 tensor cores.
"""
import ctypes

import numpy as np
import torch

from ..core.util import assure_2d
from .. import _lib
from ..core.integration.engine import uniform_rvs


def _dev(x):
    return _lib.to_device(x)[0]


class FunctionBasis(object):
    mode = None

    def __init__(self, dim):
        self.dim = dim

    # ---- parameters -> device block ---------------------------------------------
    def _tables(self, params):
        raise NotImplementedError

    def _block(self, params, out_bias=0.0, out_weights=None):
        p0, p1 = self._tables(params)
        ow = _dev(np.zeros(p0.shape[0]) if out_weights is None else out_weights).reshape(-1)
        blk = _lib.ElmBlock()
        blk.mode = self.mode
        blk.nodes = int(p0.shape[0])
        blk.p0_dev = p0.data_ptr()
        blk.p1_dev = p1.data_ptr()
        blk.out_w_dev = ow.data_ptr()
        blk.out_bias = float(out_bias)
        return blk, (p0, p1, ow)

    def _run(self, what, params, out_bias, out_weights, xs, precision=0):
        x_dev, host = _lib.to_device(xs)
        if x_dev.dim() < 2:
            x_dev = x_dev.reshape(1, -1) if self.dim > 1 or x_dev.dim() == 0 else x_dev.reshape(-1, 1)
        n = x_dev.shape[0]
        blk, keep = self._block(params, out_bias, out_weights)
        shape = {0: (n,), 1: (n, self.dim), 2: (n, blk.nodes)}[what]
        out = torch.empty(shape, dtype=torch.float64, device=x_dev.device)
        _lib.call("hepmc_elm_eval", ctypes.byref(blk), self.dim, what, precision, _lib.ptr(x_dev), n,
                  _lib.ptr(out), _lib.stream_ptr())
        del keep
        return _lib.to_host_like(out, host)

    # ---- reference API ------------------------------------------------------------
    def output_matrix(self, xs, params):
        """(N, node_count) hidden-layer outputs (extreme_learning.py:120-129,190-207)."""
        return self._run(2, params, 0.0, None, assure_2d(xs))

    def random_params(self, node_count):
        raise NotImplementedError

    def eval(self, params, out_bias, out_weights, xs, precision=0):
        """Surrogate value at xs (extreme_learning.py:18-29)."""
        return self._run(0, params, out_bias, out_weights, xs, precision)

    def eval_gradient(self, params, out_bias, out_weights, xs, precision=0):
        """Surrogate gradient at xs -> (N, dim) (extreme_learning.py:140-146,222-230)."""
        return self._run(1, params, out_bias, out_weights, xs, precision)

    def _split(self, xs):
        if np.isscalar(xs[0]):
            return np.stack(xs, axis=0), None
        shape = np.asanyarray(xs[0]).shape
        return np.stack([np.asanyarray(x).flatten() for x in xs], axis=1), shape

    def eval_split(self, params, out_bias, out_weights, *xs):
        pts, shape = self._split(xs)
        res = self.eval(params, out_bias, out_weights, pts)
        return res if shape is None else res.reshape(shape)

    def eval_gradient_split(self, params, out_bias, out_weights, *xs):
        pts, shape = self._split(xs)
        res = self.eval_gradient(params, out_bias, out_weights, pts)
        return res if shape is None else res.reshape((*shape, self.dim))

    def extreme_learning_train(self, xs, values, node_count, params=None, ridge=0.0):
        """Least-squares output weights for random hidden parameters (:56-67): design
        matrix from the output_matrix kernel, pseudo-inverse on the device
        (torch.linalg.pinv -> cuSOLVER; a library call, not a hot-path kernel).

        :param ridge: extension over the reference (0 = its plain pseudo-inverse).  A relative
            Tikhonov term lambda = ridge * trace(A^T A) / nodes.  With many overlapping nodes the
            plain pseudo-inverse returns huge weights of alternating sign (|w| ~ 1e7 for 1024
            nodes); ridge = 1e-6 keeps the fit (rms 0.06 vs 0.02 on the 8-D camel) and the HMC
            acceptance, and brings the cancellation in the gradient sum from ~1e4 down to ~60 --
            which is what the TF32 tensor-core gradient needs (SurrogateGradient.cancellation)."""
        xs_dev, host = _lib.to_device(assure_2d(xs))
        if params is None:
            params = self.random_params(node_count)
        v = _dev(values).reshape(-1)
        out_bias = float(v.mean().item())
        design = self._run(2, params, 0.0, None, xs_dev)
        if ridge:
            gram = design.T @ design
            lam = float(ridge) * float(torch.trace(gram)) / gram.shape[0]
            gram.diagonal().add_(lam)
            weights = torch.linalg.solve(gram, design.T @ (v - out_bias))
        else:
            weights = torch.linalg.pinv(design) @ (v - out_bias)
        return params, out_bias, (weights.cpu().numpy() if host else weights).flatten()


class AdditiveBasis(FunctionBasis):
    """g(b_i + W_i . x) bases (extreme_learning.py:70-146)."""

    def __init__(self, dim, weight_range=(-1, 1), bias_range=(0, -1)):
        super().__init__(dim)
        self.weight_min = weight_range[0]
        self.weight_delta = weight_range[1] - weight_range[0]
        self.bias_min = bias_range[0]
        self.bias_delta = bias_range[1] - bias_range[0]

    def random_fn_params(self, node_count):
        return None

    def _tables(self, params):
        biases, in_weights, _ = params
        return _dev(in_weights).reshape(-1, self.dim).contiguous(), _dev(biases).reshape(-1)

    def random_params(self, node_count):
        u = uniform_rvs(node_count, self.dim + 1).cpu().numpy()
        biases = self.bias_min + self.bias_delta * u[:, 0]
        input_weights = self.weight_min + self.weight_delta * u[:, 1:]
        return biases, input_weights, self.random_fn_params(node_count)


class RadialBasis(FunctionBasis):
    """g(|x - c_i|^2 / (2 w_i^2)) bases with scalar widths (extreme_learning.py:149-230)."""

    def __init__(self, dim, width_range=(.01, 1), center_range=(0, 1), multi_widths=False):
        super().__init__(dim)
        if multi_widths:
            raise NotImplementedError("multi_widths is unusable with eval_gradient in the reference "
                                      "(shape error, SURVEY.md App. A.9) and is not supported")
        self.weight_min = width_range[0]
        self.weight_delta = width_range[1] - width_range[0]
        self.bias_min = center_range[0]
        self.bias_delta = center_range[1] - center_range[0]
        self.multi_width = False

    def random_fn_params(self, node_count):
        return None

    def _tables(self, params):
        centers, widths, _ = params
        return _dev(centers).reshape(-1, self.dim).contiguous(), _dev(widths).reshape(-1)

    def random_params(self, node_count):
        u = uniform_rvs(node_count, self.dim + 1).cpu().numpy()
        centers = self.bias_min + self.bias_delta * u[:, :self.dim]
        widths = self.weight_min + self.weight_delta * u[:, self.dim]
        return centers, widths, self.random_fn_params(node_count)


class TrigBasis(AdditiveBasis):
    mode = _lib.GRAD_ELM_TRIG

    def __init__(self, dim, weight_range=(0, 1), bias_range=(-1, 0)):
        super().__init__(dim, weight_range, bias_range)


class GaussianBasis(RadialBasis):
    mode = _lib.GRAD_ELM_GAUSS

    def __init__(self, ndim, width_range=(.01, 1), center_range=(0, 1), multi_widths=False):
        super().__init__(ndim, width_range, center_range, multi_widths)


class SurrogateGradient(object):
    """The callable the reference wires into `target.pot_gradient`
    (interfaces/mc3_hmc_el.py:32-35: `lambda xs: basis.eval_gradient(*params, xs)`), as an
    object so HamiltonianUpdate can hand its node tables to the chain kernel."""

    def __init__(self, basis, params, out_bias, out_weights):
        self.basis = basis
        self.params = params
        self.out_bias = out_bias
        self.out_weights = out_weights

    def __call__(self, xs):
        return self.basis.eval_gradient(self.params, self.out_bias, self.out_weights, xs)

    def block(self):
        return self.basis._block(self.params, self.out_bias, self.out_weights)

    def cancellation(self, xs):
        """Median over points and coordinates of  sum_i |term_i| / |sum_i term_i|  for the
        Gaussian-basis gradient (term_i = basis_i(x) * out_w_i / width_i^2 * (x - c_i)).
        The tensor-core path keeps TF32 operands in its second GEMM: every term carries a
        relative error of ~2^-11, so the gradient's relative error is about that times this
        ratio.  Measured on the 8-D camel: ratios up to ~300 leave the HMC acceptance unchanged,
        1e4 (plain pseudo-inverse, 1024 nodes) lowers it from 0.83 to 0.4-0.6."""
        if self.basis.mode != _lib.GRAD_ELM_GAUSS:
            raise TypeError("cancellation() is defined for the Gaussian basis (the tensor-core path)")
        x, _ = _lib.to_device(assure_2d(xs))
        centers = _dev(self.params[0])
        widths = _dev(self.params[1]).reshape(-1)
        coef = _dev(self.out_weights).reshape(-1) / widths ** 2
        diff = x[:, None, :] - centers[None, :, :]
        p = torch.exp(-(diff ** 2).sum(dim=2) / (2 * widths ** 2))
        terms = (p * coef)[:, :, None] * diff
        ratio = terms.abs().sum(dim=1) / terms.sum(dim=1).abs().clamp_min(1e-300)
        return float(ratio.median())


def attach_surrogate_gradient(target, basis, params, out_bias, out_weights):
    """target.pot_gradient = surrogate gradient; returns the SurrogateGradient."""
    grad = SurrogateGradient(basis, params, out_bias, out_weights)
    target.pot_gradient = grad
    return grad
