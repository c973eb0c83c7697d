"""Explicit counter-based RNG state: (seed, stream id).

The reference draws from NumPy's global MT19937 (`np.random.*`); the kernels use
Philox4x32-10 with key = seed and counter = (item index, draw block, stream id).
Every consuming call takes a fresh stream id, so successive calls are independent
and a run is reproducible from `seed()` alone -- at ANY GPU count, because the
counter uses global item indices.
"""
_state = {"seed": 1234, "stream": 0}


def seed(value=1234):
    """Counterpart of np.random.seed for the native generator."""
    _state["seed"] = int(value) & 0xFFFFFFFFFFFFFFFF
    _state["stream"] = 0


def get_seed():
    return _state["seed"]


def next_stream(count=1):
    """Reserve `count` consecutive stream ids; returns the first."""
    s = _state["stream"]
    _state["stream"] = (s + int(count)) & 0xFFFFFFFF
    return s


def get_state():
    return dict(_state)


def set_state(state):
    _state.update(state)
