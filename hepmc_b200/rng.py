"""Explicit counter-based RNG state: (seed, stream id).

The reference draws from NumPy's global MT19937 (`np.random.*`); the kernels use
Philox4x32-10 with key = seed and counter = (item index, draw block, stream id).
Every consuming call takes a fresh stream id, so successive calls are independent
and a run is reproducible from `seed()` alone -- at ANY GPU count, because the
counter uses global item indices.
"""
_state = {"seed": 1234, "stream": 0}

# Draw modes of the integrators' native generator (include/hepmc_b200.h, enum hepmc_rng_mode):
# bits per uniform / per VEGAS (bin, offset) pair and Philox4x32 round count.
#   "u32r10" (default): 32 random bits per axis, Philox4x32-10 -- uniform (w + 1/2) 2^-32; VEGAS
#            t = K w: bin = t >> 32, offset = (t mod 2^32 + 1/2) 2^-32; four axes per Philox call
#   "u52r10": 52-bit uniforms / 64 random bits per (bin, offset) pair, two axes per call
#   "u32r7":  as u32r10 with Philox4x32-7 (the Crush-resistant minimum of Salmon et al.)
# The mode applies to PlainMC / VegasMC / StratifiedMC / GridVolumes draws and to phase_space.Rambo.generate.
# The chain kernels and every helper that returns raw uniforms always draw 52-bit uniforms with 10 rounds.
MODES = {"u52r10": 0, "u32r10": 1, "u32r7": 2}
DEFAULT_MODE = "u32r10"
_state["mode"] = DEFAULT_MODE


def set_mode(name=DEFAULT_MODE):
    """Select the draw mode of the integrators' native draws (this process)."""
    if name not in MODES:
        raise ValueError("unknown RNG mode %r (one of %s)" % (name, ", ".join(sorted(MODES))))
    _state["mode"] = name


def get_mode():
    return _state["mode"]


def mode_code():
    """The C ABI's rng_mode argument for the current mode."""
    return MODES[_state["mode"]]


def seed(value=1234):
    """Counterpart of np.random.seed for the native generator."""
    _state["seed"] = int(value) & 0xFFFFFFFFFFFFFFFF
    _state["stream"] = 0     # the draw mode is a setting, not part of the seed state


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
