"""Launch helpers for the chain-ensemble kernels (hepmc_rwm_chains, hepmc_hmc_chains).

An ensemble is `C` independent chains identified by GLOBAL chain indices
`first_chain .. first_chain + C - 1`; under torch.distributed each rank runs a contiguous
index range with no collective on the path (acceptance totals are summed on request).
"""
import ctypes

import numpy as np
import torch

from ... import _lib, rng


def gaussian_rvs(mean, scale, n):
    """(n, ndim) draws mean + scale * z, z standard normal from the native generator."""
    from ..integration.engine import uniform_rvs
    ndim = len(mean)
    npair = (ndim + 1) // 2
    u = uniform_rvs(n, 2 * npair)
    r = torch.sqrt(-2.0 * torch.log(u[:, 0::2]))
    ang = 2.0 * np.pi * u[:, 1::2]
    z = torch.stack((r * torch.cos(ang), r * torch.sin(ang)), dim=2).reshape(n, 2 * npair)[:, :ndim]
    dev = z.device
    return (torch.as_tensor(np.asarray(mean, dtype=np.float64), device=dev) +
            torch.as_tensor(np.asarray(scale, dtype=np.float64), device=dev) * z).contiguous()


def prepare_initial(initial, ndim):
    """initial -> ((C, ndim) CUDA tensor, is_ensemble).  A 1-D `initial` of length ndim is a
    single chain (the reference's only mode, base.py:56-59); a 2-D (C, ndim) array starts C
    independent chains (the signature-preserving ensemble extension)."""
    if isinstance(initial, torch.Tensor):
        arr = initial
        nd = arr.dim()
    else:
        arr = np.atleast_1d(np.asarray(initial, dtype=np.float64))
        nd = arr.ndim
    if nd == 1:
        if arr.shape[0] != ndim:
            raise ValueError('initial must have dimension ' + str(ndim))
        x0, _ = _lib.to_device(arr.reshape(1, ndim))
        return x0, False
    if nd == 2:
        if arr.shape[1] != ndim:
            raise ValueError('initial must have dimension ' + str(ndim))
        x0, _ = _lib.to_device(arr)
        return x0, True
    raise ValueError('initial must be (ndim,) or (n_chains, ndim)')


def alloc_outputs(n_chains, ndim, sample_size, thin, store_chain, dev):
    n_kept = (sample_size + thin - 1) // thin
    chain = torch.empty((n_chains, n_kept, ndim), dtype=torch.float64, device=dev) if store_chain else None
    last = torch.empty((n_chains, ndim), dtype=torch.float64, device=dev)
    accepted = torch.zeros(n_chains, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    return chain, last, accepted, total


def run_rwm(blk, x0, proposal_kind, prop_param, sample_size, thin=1, store_chain=True, first_chain=0,
            stream_id=None, replay=None):
    dev = x0.device
    n_chains, ndim = x0.shape
    chain, last, accepted, total = alloc_outputs(n_chains, ndim, sample_size, thin, store_chain, dev)
    z = u = None
    if replay is not None:
        z, _ = _lib.to_device(replay[0])
        u, _ = _lib.to_device(replay[1])
        if tuple(z.shape) != (n_chains, sample_size - 1, ndim) or tuple(u.shape) != (n_chains, sample_size - 1):
            raise ValueError("replay tapes must be (C, T-1, ndim) and (C, T-1)")
    sid = rng.next_stream() if stream_id is None else stream_id
    _lib.call("hepmc_rwm_chains", ctypes.byref(blk), _lib.ptr(x0), int(proposal_kind), _lib.dbl_array(prop_param),
              n_chains, int(first_chain), int(sample_size), int(thin), rng.get_seed(), sid,
              _lib.ptr(z), _lib.ptr(u), _lib.ptr(chain), _lib.ptr(last), _lib.ptr(accepted), _lib.ptr(total),
              _lib.stream_ptr())
    return chain, last, accepted, total


def coop_flag(n_local, nodes, total_chains=None):
    """The `coop` argument of hepmc_hmc_chains / hepmc_mc3_chains / hepmc_leapfrog (1 = one CTA per chain,
    0 = one chain per thread), decided from the GLOBAL ensemble size so that the same chain runs the
    same kernel -- hence the same arithmetic -- however the ensemble is sharded over calls or GPUs:
    `total_chains` if the caller states it, else the sum of the local counts over the process group,
    else the local count.  Threshold: csrc/chains.cuh coop_max_chains (8 * nodes in [1024, 8192])."""
    if total_chains is None:
        from ... import parallel
        rank, ws = parallel.world()
        total_chains = n_local
        if ws > 1:
            t = torch.tensor([n_local], dtype=torch.int64, device=_lib.device())
            total_chains = int(parallel.sum_int(t).item())
    limit = min(max(8 * int(nodes), 1024), 8192)
    return 1 if int(total_chains) <= limit else 0


def run_hmc(blk, elm_blk, precision, x0, mass_diag, step_size, steps, sample_size, thin=1, store_chain=True,
            first_chain=0, stream_id=None, replay=None, total_chains=None):
    dev = x0.device
    n_chains, ndim = x0.shape
    chain, last, accepted, total = alloc_outputs(n_chains, ndim, sample_size, thin, store_chain, dev)
    p = u = None
    if replay is not None:
        p, _ = _lib.to_device(replay[0])
        u, _ = _lib.to_device(replay[1])
        if tuple(p.shape) != (n_chains, sample_size - 1, ndim) or tuple(u.shape) != (n_chains, sample_size - 1):
            raise ValueError("replay tapes must be (C, T-1, ndim) and (C, T-1)")
    sid = rng.next_stream() if stream_id is None else stream_id
    coop = coop_flag(n_chains, elm_blk.nodes, total_chains) if elm_blk is not None else 0
    _lib.call("hepmc_hmc_chains", ctypes.byref(blk), ctypes.byref(elm_blk) if elm_blk is not None else None,
              int(precision), coop, _lib.ptr(x0), _lib.dbl_array(mass_diag), float(step_size), int(steps),
              n_chains, int(first_chain), int(sample_size), int(thin), rng.get_seed(), sid,
              _lib.ptr(p), _lib.ptr(u), _lib.ptr(chain), _lib.ptr(last), _lib.ptr(accepted), _lib.ptr(total),
              _lib.stream_ptr())
    return chain, last, accepted, total


# ---------------------------------------------------------------------------------------------
# Generic path: targets that are NOT built-in densities (user classes / callables).  The chain
# kernels cannot call Python, so -- like the integrators' two-kernel path for user integrands --
# every step is: draws from the native generator (one kernel, Philox counter = global chain
# index, one stream id per step) -> proposal arithmetic -> the USER's pdf / pot_gradient on the
# (C, ndim) CUDA tensor -> accept / reject, all on the device, vectorised over the ensemble.
# Cost is a few dozen launches per step whatever the ensemble size: for one chain this is on par
# with the reference, for 1e5 chains it is 1e5 times that.  No host arithmetic, no CPU fallback.
# ---------------------------------------------------------------------------------------------
def _user_values(fn, x, what):
    """fn(x) for x (C, ndim) on the device -> float64 CUDA tensor; fn may answer with a tensor or
    anything np.asarray understands (util.host_function marks NumPy code: it gets a host copy)."""
    from ..util.helper import call_user_function
    out = call_user_function(fn, [x], "the target's " + what)
    if not isinstance(out, torch.Tensor):
        out = torch.as_tensor(np.asarray(out, dtype=np.float64))
    out = out.to(device=x.device, dtype=torch.float64)
    if out.dim() == 0:                       # a constant (`lambda x: 1`, reference tests/test_markov.py:10)
        out = out.expand(x.shape[0])
    return out


def _step_draws(n_chains, n_uniform, first_chain, stream_id, dev):
    """(C, n_uniform) uniforms of one step; chain c uses counter first_chain + c."""
    from ..integration.engine import none_block
    out = torch.empty((n_chains, n_uniform), dtype=torch.float64, device=dev)
    blk = none_block(n_uniform)
    _lib.call("hepmc_plain_mc_sample", ctypes.byref(blk), n_chains, int(first_chain), _lib.default_chunk(n_chains),
              rng.get_seed(), int(stream_id), rng.MODES["u52r10"], None, _lib.ptr(out), None, None, _lib.stream_ptr())
    return out


def _normals(u, ndim):
    """Box-Muller on column pairs of u -> (C, ndim) standard normals."""
    npair = (ndim + 1) // 2
    r = torch.sqrt(-2.0 * torch.log(u[:, 0:2 * npair:2]))
    ang = 2.0 * np.pi * u[:, 1:2 * npair:2]
    return torch.stack((r * torch.cos(ang), r * torch.sin(ang)), dim=2).reshape(u.shape[0], 2 * npair)[:, :ndim]


def _metropolis_accept(state, pdf, cand, cpdf, prob, u, accepted):
    take = (prob >= 1) | (u < prob)                       # metropolis.py:87; NaN -> reject
    changed = take & (cand != state).any(dim=1)           # base.py:72-73
    accepted += changed.to(torch.int64)
    return torch.where(take[:, None], cand, state), torch.where(take, cpdf, pdf)


def run_rwm_generic(pdf_fn, x0, proposal_kind, prop_param, sample_size, thin=1, store_chain=True, first_chain=0):
    """DefaultMetropolis over a user-defined target (symmetric proposals: kinds 0, 1, 2 of
    hepmc_rwm_chains)."""
    dev = x0.device
    C, ndim = x0.shape
    chain, last, accepted, total = alloc_outputs(C, ndim, sample_size, thin, store_chain, dev)
    par = torch.as_tensor(np.array(prop_param, dtype=np.float64), device=dev)
    npair = (ndim + 1) // 2
    n_draw = 2 * npair if proposal_kind == 0 else (ndim if proposal_kind == 1 else 1)
    sid0 = rng.next_stream(max(sample_size - 1, 1))
    state = x0.clone()
    pdf = _user_values(pdf_fn, state, "pdf").reshape(C)
    if chain is not None:
        chain[:, 0] = state
    for t in range(1, sample_size):
        u = _step_draws(C, n_draw + 1, first_chain, sid0 + t - 1, dev)
        if proposal_kind == 0:
            cand = state + par[:ndim] * _normals(u, ndim)
        elif proposal_kind == 1:
            cand = par[0] + (par[1] - par[0]) * u[:, :ndim]
        else:
            delta, lo, hi = par[:ndim], par[ndim], par[ndim + 1]
            cand = torch.minimum(torch.clamp(state - delta / 2, min=float(lo)), hi - delta) + u[:, :1] * delta
        cpdf = _user_values(pdf_fn, cand, "pdf").reshape(C)
        state, pdf = _metropolis_accept(state, pdf, cand, cpdf, cpdf / pdf, u[:, -1], accepted)
        if chain is not None and t % thin == 0:
            chain[:, t // thin] = state
    last.copy_(state)
    total += accepted.sum()
    return chain, last, accepted, total


def run_hmc_generic(pdf_fn, grad_fn, x0, mass_diag, step_size, steps, sample_size, thin=1, store_chain=True,
                    first_chain=0, simulate=None):
    """HamiltonianUpdate over a user-defined target: leapfrog of simulation.py:34-42 with the user's
    pot_gradient, momentum ~ N(0, diag(mass)), acceptance of hmc.py:54-59.  `simulate`: a user-supplied
    simulator (hmc.py:25-27 `simulate=` / `sim_class=`), called as simulate(q, p) on (C, ndim) CUDA
    tensors instead of the built-in loop; rows it returns non-finite are rejected (hmc.py:77-78)."""
    dev = x0.device
    C, ndim = x0.shape
    chain, last, accepted, total = alloc_outputs(C, ndim, sample_size, thin, store_chain, dev)
    mass = torch.as_tensor(np.array(mass_diag, dtype=np.float64), device=dev)
    npair = (ndim + 1) // 2
    sid0 = rng.next_stream(max(sample_size - 1, 1))
    state = x0.clone()
    pdf = _user_values(pdf_fn, state, "pdf").reshape(C)
    if chain is not None:
        chain[:, 0] = state
    half = step_size / 2 if step_size is not None else None

    def grad(q):
        return _user_values(grad_fn, q, "pot_gradient").reshape(C, ndim)

    for t in range(1, sample_size):
        u = _step_draws(C, 2 * npair + 1, first_chain, sid0 + t - 1, dev)
        p0 = torch.sqrt(mass) * _normals(u, ndim)
        q, p = state.clone(), p0.clone()
        if simulate is not None:
            out = simulate(q, p)
            if out is None or out[0] is None:
                q, p = torch.full_like(state, float("nan")), torch.full_like(state, float("nan"))
            else:
                q = _user_values(lambda _x: out[0], q, "simulate").reshape(C, ndim)
                p = _user_values(lambda _x: out[1], p, "simulate").reshape(C, ndim)
        else:
            g = grad(q)
            for _ in range(int(steps)):
                p = p - half * g
                q = q + step_size * (p / mass)
                g = grad(q)
                p = p - half * g
        cpdf = _user_values(pdf_fn, q, "pdf").reshape(C)
        if simulate is not None:
            cpdf = torch.where(torch.isfinite(q).all(dim=1), cpdf, torch.zeros_like(cpdf))   # q is None -> pdf = 0
        # N(p) / N(p0) for the diagonal Gaussian momentum distribution
        kin = torch.exp(-0.5 * ((p * p - p0 * p0) / mass).sum(dim=1))
        prob = cpdf * kin / pdf
        prob = torch.where(torch.isnan(prob), torch.zeros_like(prob), prob)      # hmc.py:58-59
        state, pdf = _metropolis_accept(state, pdf, q, cpdf, prob, u[:, -1], accepted)
        if chain is not None and t % thin == 0:
            chain[:, t // thin] = state
    last.copy_(state)
    total += accepted.sum()
    return chain, last, accepted, total


def run_mixing_generic(pdf_fn, channels, beta, local, x0, sample_size, thin=1, store_chain=True, first_chain=0):
    """MixingMarkovUpdate([IS-Metropolis over `channels`, local update]) for a user-defined target
    (base.py:169-179 + metropolis.py:37-50): per step every chain picks the importance-sampling update
    with probability beta, else the local one (DefaultMetropolis with a Gaussian / UniformLocal
    proposal, or HamiltonianUpdate with the user's pot_gradient).  Both candidates are produced for the
    whole ensemble (kernels), the user's pdf is called ONCE per step on the selected candidates."""
    from ..proposals import Gaussian as GaussianProposal, UniformLocal
    from .metropolis import DefaultMetropolis
    from ..integration.engine import none_block
    dev = x0.device
    C, ndim = x0.shape
    chain, last, accepted, total = alloc_outputs(C, ndim, sample_size, thin, store_chain, dev)
    is_hmc = not isinstance(local, DefaultMetropolis) and local is not None
    if local is None:
        if beta < 1.0:
            raise TypeError("a local update is needed when beta < 1")
    elif is_hmc:
        tgt = local.target_density
        if not callable(getattr(tgt, "pot_gradient", None)):
            raise TypeError("the local HamiltonianUpdate needs a target with a callable pot_gradient")
        mass = torch.as_tensor(np.array(local.p_dist._diag(), dtype=np.float64), device=dev)
        step_size, steps = float(local.step_size), int(local.steps)
    elif isinstance(local._proposal, GaussianProposal):
        cov = np.asarray(local._proposal.cov)
        if np.count_nonzero(cov - np.diag(np.diagonal(cov))):
            raise TypeError("the local random walk needs a diagonal proposal covariance")
        scale = torch.as_tensor(np.sqrt(np.diagonal(cov)), device=dev)
    elif isinstance(local._proposal, UniformLocal):
        par = torch.as_tensor(np.array(local._proposal._kernel_params(), dtype=np.float64), device=dev)
    else:
        raise TypeError("the local update must be HamiltonianUpdate or DefaultMetropolis with proposals.Gaussian / "
                        "proposals.UniformLocal, got " + type(local).__name__)
    npair = (ndim + 1) // 2
    sid0 = rng.next_stream(2 * max(sample_size - 1, 1))        # per step: one stream for the mixture draw, one for the rest
    table = channels._table()
    blk = none_block(ndim)
    chunk = _lib.default_chunk(C)
    state = x0.clone()
    pdf = _user_values(pdf_fn, state, "pdf").reshape(C)
    if chain is not None:
        chain[:, 0] = state
    is_cand = torch.empty((C, ndim), dtype=torch.float64, device=dev)
    q_cand = torch.empty(C, dtype=torch.float64, device=dev)
    ch = torch.empty(C, dtype=torch.int64, device=dev)
    for t in range(1, sample_size):
        u = _step_draws(C, 2 * npair + 2, first_chain, sid0 + 2 * (t - 1), dev)
        pick_is = u[:, -2] < beta                                            # base.py:170
        # importance-sampling candidate ~ mixture, with q(candidate) from the same kernel
        _lib.call("hepmc_multichannel_sample", ctypes.byref(blk), _lib.ptr(table), channels.count, ndim,
                  C, int(first_chain), chunk, rng.get_seed(), sid0 + 2 * (t - 1) + 1, None, None,
                  _lib.ptr(is_cand), None, _lib.ptr(q_cand), _lib.ptr(ch), None, None, _lib.stream_ptr())
        if local is None:
            loc_cand, kin = is_cand, None
        elif is_hmc:
            p0 = torch.sqrt(mass) * _normals(u, ndim)
            q, p = state.clone(), p0.clone()
            g = _user_values(tgt.pot_gradient, q, "pot_gradient").reshape(C, ndim)
            for _ in range(steps):
                p = p - (step_size / 2) * g
                q = q + step_size * (p / mass)
                g = _user_values(tgt.pot_gradient, q, "pot_gradient").reshape(C, ndim)
                p = p - (step_size / 2) * g
            loc_cand = q
            kin = torch.exp(-0.5 * ((p * p - p0 * p0) / mass).sum(dim=1))
        elif isinstance(local._proposal, GaussianProposal):
            loc_cand, kin = state + scale * _normals(u, ndim), None
        else:
            delta, lo, hi = par[:ndim], par[ndim], par[ndim + 1]
            loc_cand = torch.minimum(torch.clamp(state - delta / 2, min=float(lo)), hi - delta) + u[:, :1] * delta
            kin = None
        cand = torch.where(pick_is[:, None], is_cand, loc_cand)
        cpdf = _user_values(pdf_fn, cand, "pdf").reshape(C)
        q_state = channels.pdf(state).reshape(C)
        prob_is = cpdf * q_state / pdf / q_cand                              # metropolis.py:46-47
        prob_loc = cpdf / pdf if kin is None else cpdf * kin / pdf
        prob = torch.where(pick_is, prob_is, prob_loc)
        if kin is not None:
            prob = torch.where(torch.isnan(prob) & ~pick_is, torch.zeros_like(prob), prob)   # hmc.py:58-59
        state, pdf = _metropolis_accept(state, pdf, cand, cpdf, prob, u[:, -1], accepted)
        if chain is not None and t % thin == 0:
            chain[:, t // thin] = state
    last.copy_(state)
    total += accepted.sum()
    return chain, last, accepted, total
