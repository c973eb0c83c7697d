"""Base contracts: Proposal, Density, Distribution (mirror of hepmc/core/density.py).

Signatures are the reference's: `pdf(xs)`, `pot(xs)`, `pot_gradient(xs)`,
`pdf_gradient(xs)`, `__call__(*xs)`, `rvs(n)`, `proposal(state)`,
`proposal_pdf(state, candidate)`.  Arrays may be NumPy arrays (results come back as
NumPy arrays) or torch CUDA tensors (results stay on the device).

Built-in analytic densities (camel, banana, gaussian, uniform) derive from
`BuiltinDensity`: their methods launch `hepmc_density_eval` and the integrators /
samplers recognise them and fuse them into their kernels.  User-defined densities
(`Density.make`) keep running the user's own callables; the generic `pot` /
`pot_gradient` rules of density.py:48-68 are applied to whatever they return.
"""
import ctypes

import numpy as np
import torch

from .util import interpret_array
from .. import _lib


class _AnyDensity(object):
    def __init__(self, ndim, is_symmetric=False):
        self.ndim = ndim
        self.is_symmetric = is_symmetric


class Proposal(_AnyDensity):

    def proposal(self, state):
        raise NotImplementedError()

    def proposal_pdf(self, state, candidate):
        return None  # symmetric proposals need no density

    @classmethod
    def make(cls, proposal, ndim=None, proposal_pdf=None, symmetric=None):
        if isinstance(proposal, Proposal):
            return proposal
        if ndim is None:
            raise RuntimeError("If proposal is a function, ndim must be given.")
        if symmetric is None:
            symmetric = proposal_pdf is None
        obj = cls(ndim, symmetric)
        obj.proposal = proposal
        if proposal_pdf is not None:
            obj.proposal_pdf = proposal_pdf
        return obj


def _is_tensor(x):
    return isinstance(x, torch.Tensor)


class Density(_AnyDensity):

    def __call__(self, *xs):
        """Integrand adaptor (density.py:39-46): fn(x1, .., xn) with equal-shape arrays."""
        if np.isscalar(xs[0]):
            return self.pdf(np.stack(xs, axis=0))
        if any(_is_tensor(x) for x in xs):
            shape = xs[0].shape
            cols = [torch.as_tensor(x).reshape(-1) for x in xs]
            return self.pdf(torch.stack(cols, dim=1)).reshape(shape)
        shape = np.asanyarray(xs[0]).shape
        cols = [np.asanyarray(x).flatten() for x in xs]
        return self.pdf(np.stack(cols, axis=1)).reshape(shape)

    def pot(self, xs):
        """-log pdf with NaN -> inf (density.py:48-53)."""
        pdf = self.pdf(xs)
        if _is_tensor(pdf):
            pot = -torch.log(pdf)
            return torch.where(torch.isnan(pot), torch.full_like(pot, float("inf")), pot)
        with np.errstate(divide='ignore'):
            pot = -np.log(pdf)
        pot[np.isnan(pot)] = np.inf
        return pot

    def pot_gradient(self, xs):
        """-pdf_gradient/pdf where pdf != 0, inf elsewhere (density.py:55-68)."""
        xs = interpret_array(xs, self.ndim)
        pdf = self.pdf(xs)
        if _is_tensor(xs):
            nonzero = (pdf != 0).reshape(-1)
            res = torch.full(xs.shape, float("inf"), dtype=xs.dtype, device=xs.device)
            if bool(nonzero.any()):
                res[nonzero] = -self.pdf_gradient(xs[nonzero]) / pdf[nonzero][:, None]
            return res
        pdf = np.asarray(pdf)[:, np.newaxis]
        res = np.empty(xs.shape)
        nonzero = (pdf != 0).flatten()
        res[nonzero] = -self.pdf_gradient(xs[nonzero]) / pdf[nonzero]
        res[np.logical_not(nonzero)] = np.inf
        return res

    def pdf(self, xs):
        raise NotImplementedError()

    def pdf_gradient(self, xs):
        raise NotImplementedError()

    @classmethod
    def make(cls, pdf=None, ndim=None, pdf_vect=None, is_symmetric=False,
             pdf_gradient=None, variance=None, mean=None):
        """Wrap user callables into a Density (density.py:77-99): `pdf` takes the (N, ndim) array,
        `pdf_vect` takes ndim column arrays; results are flattened to (N,).  The callables receive
        whatever array type the caller passes (NumPy arrays or CUDA tensors)."""
        if isinstance(pdf, Density):
            obj = pdf
        elif ndim is None:
            raise RuntimeError("If first argument is not a Density, ndim must be specified.")
        else:
            obj = cls(ndim, is_symmetric)

            def flat(values):
                return values.reshape(-1) if _is_tensor(values) else np.atleast_1d(values).flatten()
            if pdf is not None:
                obj.pdf = lambda xs, _f=pdf: flat(_f(xs))
            elif pdf_vect is not None:
                obj.pdf = lambda xs, _f=pdf_vect: flat(_f(*(xs.t() if _is_tensor(xs) else xs.transpose())))
        for name, value in (("pdf_gradient", pdf_gradient), ("variance", variance), ("mean", mean)):
            if value is not None:
                setattr(obj, name, value)
        return obj


class Distribution(Density, Proposal):

    def rvs(self, sample_count):
        raise NotImplementedError()

    def proposal(self, state=None):
        return self.rvs(1)[0]

    def proposal_pdf(self, state, candidate):
        return float(self.pdf(candidate))

    @classmethod
    def make(cls, pdf=None, ndim=None, rvs=None, **kwargs):
        obj = super().make(pdf=pdf, ndim=ndim, **kwargs)
        if rvs is None:
            raise RuntimeWarning("Setting rvs with None.")
        obj.rvs = lambda count: interpret_array(rvs(count), ndim)
        return obj


class BuiltinDensity(Density):
    """An analytic density the CUDA kernels know (hepmc_density_t).  Subclasses fill
    `_fill_block`; everything numerical runs in `hepmc_density_eval`."""

    kind = None

    def _fill_block(self, blk):
        raise NotImplementedError()

    def _block(self):
        if self.ndim < 1 or self.ndim > _lib.MAX_DIM:
            raise RuntimeError("built-in densities support 1 <= ndim <= %d" % _lib.MAX_DIM)
        blk = _lib.DensityBlock()
        blk.kind = self.kind
        blk.ndim = self.ndim
        self._fill_block(blk)
        return blk

    def _eval(self, what, xs):
        xs = interpret_array(xs, self.ndim)
        if xs.shape[1] != self.ndim:
            raise RuntimeWarning("Mismatching dimensions.")
        x_dev, host = _lib.to_device(xs)
        n = x_dev.shape[0]
        shape = (n,) if what in (_lib.PDF, _lib.POT) else (n, self.ndim)
        out = torch.empty(shape, dtype=torch.float64, device=x_dev.device)
        blk = self._block()
        _lib.call("hepmc_density_eval", ctypes.byref(blk), what, _lib.ptr(x_dev), n, _lib.ptr(out),
                  _lib.stream_ptr())
        return _lib.to_host_like(out, host)

    def pdf(self, xs):
        return self._eval(_lib.PDF, xs)

    def pot(self, xs):
        return self._eval(_lib.POT, xs)

    def pot_gradient(self, xs):
        return self._eval(_lib.POT_GRADIENT, xs)

    def pdf_gradient(self, xs):
        return self._eval(_lib.PDF_GRADIENT, xs)


def as_builtin(fn):
    """The built-in density behind an integrand/target, or None.

    Accepts the object itself or its bound `pdf` method; a density whose `pdf` attribute
    was replaced by the user (count_calls, monkey patches) no longer qualifies."""
    obj = fn
    if not isinstance(obj, BuiltinDensity):
        obj = getattr(fn, "__self__", None)
        if not isinstance(obj, BuiltinDensity) or getattr(fn, "__name__", "") != "pdf":
            return None
    if "pdf" in vars(obj):
        return None
    # a user subclass that overrides pdf at class level must not be fused as its analytic parent
    if not type(obj).__module__.startswith("hepmc_b200.") and type(obj).pdf is not BuiltinDensity.pdf:
        return None
    return obj
