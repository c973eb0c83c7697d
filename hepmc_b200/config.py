"""Package-level switches."""
_cfg = {"outputs": "numpy"}


def set_outputs(kind):
    """Type of arrays returned by calls that take no array input (rvs, integrator samples,
    chains): "numpy" (drop-in default: results are copied to the host like the reference's)
    or "torch" (CUDA tensors stay on the device).  Calls that take arrays always answer in
    the type they were given."""
    if kind not in ("numpy", "torch"):
        raise ValueError("outputs must be 'numpy' or 'torch'")
    _cfg["outputs"] = kind


def outputs():
    return _cfg["outputs"]


def deliver(t):
    """Hand a device tensor to the caller in the configured output type."""
    if t is None:
        return None
    return t.cpu().numpy() if _cfg["outputs"] == "numpy" else t
