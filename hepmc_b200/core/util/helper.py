"""Shape contract and small host helpers (mirror of hepmc/core/util/helper.py).

These run on the host: they decide shapes and wrap user callables; the numerical
work they used to guard now happens in the CUDA kernels.
"""
import numpy as np
import torch


def _as_array(array):
    if isinstance(array, torch.Tensor):
        return array if array.dim() >= 1 else array.reshape(1)
    return np.array(array, copy=None, subok=True, ndmin=1)


def assure_2d(array):
    """(N,) -> (N, 1); (N, ndim) unchanged; anything else is an error (helper.py:4-31)."""
    array = _as_array(array)
    if array.ndim == 2:
        return array
    if array.ndim == 1:
        return array[:, None]
    raise RuntimeError("Array must be 1 or 2 dimensional.")


def interpret_array(array, ndim=None):
    """Normalise sample arrays to (N, ndim) with the reference's rules (helper.py:34-53):
    2-D passes (width checked); 1-D with ndim == 1 is N points; 1-D of length ndim is one
    point; any other 1-D shape is a RuntimeError.  Works on NumPy arrays and torch tensors."""
    array = _as_array(array)
    if array.ndim == 2:
        if ndim and ndim != array.shape[1]:
            raise RuntimeWarning("Unexpected dimension of entries in array.")
        return array
    if ndim is None:
        return array[:, None]
    if array.ndim == 1:
        if ndim == 1:
            return array[:, None]
        if array.shape[0] == ndim:
            return array[None, :]
        raise RuntimeError("Bad array shape.")
    return None


def damped_update(old, new, damping_onset, inertia):
    """Adaptation law of VEGAS / multi-channel weights (helper.py:56-72):
    new*c/(j+1+c) + old*(j+1)/(j+1+c).  The VEGAS kernel hepmc_vegas_update_grid applies
    the same formula on the device; this host version serves user code."""
    return (new * damping_onset / (inertia + 1 + damping_onset) +
            old * (inertia + 1) / (inertia + 1 + damping_onset))


def hypercube_bounded(index, null_value=0, shape=lambda xs: xs.shape[0], self_has_ndim=False):
    """Decorator for USER functions that are non-null on the open unit cube only
    (helper.py:75-103).  Host-side (NumPy) by construction: it wraps user NumPy code."""
    def get_bounded(fn):
        def fn_bounded(*args, **kwargs):
            xs = interpret_array(args[index], args[0].ndim if self_has_ndim else None)
            xs = np.asarray(xs.cpu() if isinstance(xs, torch.Tensor) else xs)
            inside = np.all((0 < xs) * (xs < 1), axis=1)
            res = np.empty(shape(xs))
            res[inside] = fn(*args[:index], xs[inside], *args[index + 1:], **kwargs)
            res[np.logical_not(inside)] = null_value
            return res
        return fn_bounded
    return get_bounded


class Counted(object):
    """Callable wrapper counting evaluated rows (helper.py:106-118)."""

    def __init__(self, fn):
        self.fn = fn
        self.count = 0

    def __call__(self, xs, *args, **kwargs):
        nd = xs.dim() if isinstance(xs, torch.Tensor) else np.asanyarray(xs).ndim
        self.count += (xs.shape[0] if nd >= 2 else 1)
        return self.fn(xs, *args, **kwargs)


def count_calls(obj, *method_names):
    for name in method_names:
        setattr(obj, name, Counted(getattr(obj, name)))


_host_warned = set()


def call_user_function(fn, device_args, what):
    """Call USER code with CUDA tensors; if it turns out to be NumPy code (the reference's integrands and
    targets are: `lambda x: np.sin(10 * x)`) -- a TypeError / "can't convert cuda tensor" from inside it --
    call it again with host copies of the arguments, as if it were wrapped in `host_function`, and say so
    once per function.  This is the user's code running where the user wrote it; the library's own
    arithmetic (draws, maps, weights, reductions, accept/reject) stays on the device either way."""
    import warnings
    if isinstance(fn, host_function):
        return fn(*[a.cpu().numpy() for a in device_args])
    if id(fn) in _host_warned:
        return fn(*[a.cpu().numpy() for a in device_args])
    try:
        return fn(*device_args)
    except (TypeError, RuntimeError) as exc:
        msg = str(exc)
        if isinstance(exc, RuntimeError) and "numpy" not in msg.lower() and "cuda" not in msg.lower():
            raise
        try:
            out = fn(*[a.cpu().numpy() for a in device_args])
        except Exception:
            raise TypeError("%s failed on CUDA tensors (%s) and on host arrays.  User functions must accept torch CUDA "
                            "tensors or NumPy arrays." % (what, exc)) from exc
        _host_warned.add(id(fn))
        warnings.warn("%s is NumPy code (%s): it is evaluated on the host from device-to-host copies of its "
                      "arguments; write it with torch operations, or wrap it in hepmc_b200.util.host_function to "
                      "silence this warning." % (what, type(exc).__name__), RuntimeWarning, stacklevel=3)
        return out


class host_function(object):
    """Marks a user integrand written for NumPy: the integrators hand it host arrays
    (device->host copy of the sample, host->device copy of the values) instead of CUDA
    tensors.  This is user code running where the user wrote it -- not a fallback of the
    library's own path."""

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, *xs):
        return self.fn(*xs)
