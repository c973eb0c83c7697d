"""Leapfrog integrator (mirror of hepmc/hamiltonian/simulation.py:4-46).

`HamiltonLeapfrog(pot_gradient, kin_gradient, step_size, steps)(q_init, p_init)` keeps the
reference's signature and loop

    g = dU(q);  steps x { p -= eps/2 g;  q += eps dK(p);  g = dU(q);  p -= eps/2 g }

(half steps not merged, steps + 1 gradient evaluations) and runs it on the device for one state
`(ndim,)` -- the reference's use -- or a batch `(C, ndim)`:

* `pot_gradient` = the `pot_gradient` of a built-in density (bound method) or a
  `surrogate.SurrogateGradient`, `kin_gradient` = `densities.Gaussian(...).pot_gradient` with a
  diagonal covariance and zero mean: ONE kernel (C ABI `hepmc_leapfrog`), the same device function
  the HMC ensemble kernel integrates with;
* any other callables (taking and returning `(C, ndim)` CUDA tensors, or `util.host_function`
  wrappers): the same loop as device tensor operations, one pass of the user's functions per step.

There is no host arithmetic and no CPU fallback.  Where the reference returns `(None, None)` on a
NumPy RuntimeWarning (overflow), a single state whose result is not finite returns `(None, None)`;
in a batch the non-finite rows are returned as they are.
"""
import ctypes

import numpy as np
import torch

from .. import _lib


def _kernel_pieces(pot_gradient, kin_gradient):
    """(density, elm block | None, keep-alive, mass diag, precision) if the pair can run in the kernel."""
    from ..core.density import BuiltinDensity
    from ..core.densities import Gaussian
    from ..surrogate.extreme_learning import SurrogateGradient
    kin_owner = getattr(kin_gradient, "__self__", None)
    if not (isinstance(kin_owner, Gaussian) and getattr(kin_gradient, "__func__", None) is Gaussian.pot_gradient):
        return None
    if np.any(np.asarray(kin_owner.mean) != 0):
        return None
    try:
        mass = kin_owner._diag()
    except Exception:
        return None
    pot = getattr(pot_gradient, "fn", pot_gradient)           # unwrap util.count_calls
    if isinstance(pot, SurrogateGradient):
        target = getattr(pot, "target", None)
        if target is None:
            # the potential's density only supplies ndim to the kernel when the gradient is a surrogate
            from ..core.densities import Gaussian as G
            target = G(len(mass))
        elm_blk, keep = pot.block()
        return target, elm_blk, keep, mass
    owner = getattr(pot, "__self__", None)
    if isinstance(owner, BuiltinDensity) and "pot_gradient" not in vars(owner) and \
            getattr(type(owner), "pot_gradient", None) is getattr(pot, "__func__", None):
        return owner, None, None, mass
    return None


class HamiltonLeapfrog(object):

    def __init__(self, pot_gradient, kin_gradient, step_size, steps, precision="f64"):
        """:param precision: arithmetic of a surrogate gradient, "f64" (default) or "f32"."""
        self.kin_gradient = kin_gradient
        self.pot_gradient = pot_gradient
        self.step_size = step_size
        self.steps = steps
        self.precision = precision

    def __call__(self, q_init, p_init, total_states=None):
        _lib.require_cuda()
        q_host = not isinstance(q_init, torch.Tensor)
        q_arr = q_init if not q_host else np.array(q_init, dtype=np.float64, ndmin=1)
        p_arr = p_init if isinstance(p_init, torch.Tensor) else np.array(p_init, dtype=np.float64, ndmin=1)
        single = q_arr.ndim == 1
        q, _ = _lib.to_device(q_arr.reshape(1, -1) if single else q_arr)
        p, _ = _lib.to_device(p_arr.reshape(1, -1) if single else p_arr)
        if q.shape != p.shape or q.dim() != 2:
            raise ValueError("q_init and p_init must both be (ndim,) or (C, ndim)")
        pieces = _kernel_pieces(self.pot_gradient, self.kin_gradient)
        if pieces is not None and q.shape[1] == pieces[0].ndim:
            q_out, p_out = self._kernel(pieces, q, p, total_states)
        else:
            q_out, p_out = self._generic(q, p)
        if single:
            if not (bool(torch.isfinite(q_out).all()) and bool(torch.isfinite(p_out).all())):
                return None, None                                   # simulation.py:43-45
            q_out, p_out = q_out[0], p_out[0]
        if q_host:
            return q_out.cpu().numpy(), p_out.cpu().numpy()
        return q_out, p_out

    def _kernel(self, pieces, q, p, total_states):
        from ..core.markov.ensemble import coop_flag
        target, elm_blk, keep, mass = pieces
        n, ndim = q.shape
        q_out, p_out = torch.empty_like(q), torch.empty_like(p)
        prec = {"f64": 0, "f32": 1}.get(self.precision)
        if prec is None:
            raise ValueError("HamiltonLeapfrog precision must be 'f64' or 'f32' (tensor-core gradients run inside "
                             "HamiltonianUpdate.sample)")
        coop = coop_flag(n, elm_blk.nodes, total_states) if elm_blk is not None else 0
        _lib.call("hepmc_leapfrog", ctypes.byref(target._block()), ctypes.byref(elm_blk) if elm_blk is not None else None,
                  prec if elm_blk is not None else 0, coop, _lib.dbl_array(mass), float(self.step_size), int(self.steps),
                  n, _lib.ptr(q), _lib.ptr(p), _lib.ptr(q_out), _lib.ptr(p_out), _lib.stream_ptr())
        del keep
        return q_out, p_out

    def _generic(self, q, p):
        """User-supplied gradients: the loop of simulation.py:36-42 on (C, ndim) device tensors."""
        from ..core.markov.ensemble import _user_values
        n, ndim = q.shape
        q, p = q.clone(), p.clone()

        def dU(x):
            return _user_values(self.pot_gradient, x, "pot_gradient").reshape(n, ndim)

        def dK(x):
            return _user_values(self.kin_gradient, x, "kin_gradient").reshape(n, ndim)

        g = dU(q)
        for _ in range(int(self.steps)):
            p = p - self.step_size / 2 * g
            q = q + self.step_size * dK(p)
            g = dU(q)
            p = p - self.step_size / 2 * g
        return q, p
