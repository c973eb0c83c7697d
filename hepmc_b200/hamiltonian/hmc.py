"""Hamiltonian Monte Carlo (mirror of hepmc/hamiltonian/hmc.py:23-109).

`HamiltonianUpdate(target_density, p_dist, steps, step_size)`: per transition draw
p ~ p_dist, leapfrog, accept with  pdf(q') N(p') / (pdf(q) N(p))  (NaN -> reject).
Device kernel requirements: `target_density` built-in (its pdf drives the Metropolis
test in float64), `p_dist` = densities.Gaussian with diagonal covariance and zero mean,
gradient = the target's own `pot_gradient` or an ELM surrogate attached with
`surrogate.attach_surrogate_gradient` / by assigning a `SurrogateGradient` to
`target_density.pot_gradient` (the wiring of interfaces/mc3_hmc_el.py:32-35).
"""
import numpy as np

from ..core.markov.metropolis import MetropolisUpdate
from ..core.markov import ensemble
from ..core.density import BuiltinDensity
from ..core.densities import Gaussian
from .simulation import HamiltonLeapfrog
from .. import _lib


# cancellation ratio of the surrogate gradient above which precision='tc' switches to 'tc32'
# (measured: <= ~300 is harmless for TF32, ~1e4 halves the acceptance rate; scripts/tc_conditioning.py)
TC_CANCELLATION_LIMIT = 2000.0


class HamiltonianUpdate(MetropolisUpdate):

    def __init__(self, target_density, p_dist, steps, step_size, simulate=None,
                 sim_class=HamiltonLeapfrog, is_adaptive=False, precision="f64"):
        """:param precision: arithmetic of a SURROGATE gradient: "f64" (parity), "f32"
        (SIMT single precision), "tc" (tcgen05 tensor cores for both GEMMs, TF32 operands in the
        contraction; switches itself to "tc32" when the trained weights cancel too heavily for
        TF32) or "tc32" (tensor cores for the exponent GEMM, FP32 contraction on the FMA pipe:
        float32-level accuracy).  The exact gradient and the Metropolis test are always float64."""
        super().__init__(target_density.ndim, target_density.pdf, is_adaptive)
        self.p_dist = p_dist
        self.target_density = target_density
        self.precision = precision
        # hmc.py:47-52.  The fused ensemble kernel integrates with HamiltonLeapfrog's loop; any OTHER simulator
        # (a `simulate` callable or a custom `sim_class`, e.g. a different splitting) is honoured by the
        # per-step device path: it is called as simulate(q, p) on (C, ndim) CUDA tensors once per transition.
        if simulate is None:
            self.simulate = sim_class(target_density.pot_gradient, p_dist.pot_gradient, step_size, steps)
        else:
            self.simulate = simulate

    def proposal_pdf(self, state, candidate):
        pass

    @property
    def precision(self):
        return self._precision

    @precision.setter
    def precision(self, value):
        if value not in ("f64", "f32", "tc", "tc32"):
            raise ValueError("precision must be 'f64', 'f32', 'tc' or 'tc32'")
        self._precision = value
        self._tc_key = None            # forget the tc / tc32 decision taken for the previous setting

    @property
    def kernel_precision(self):
        """The surrogate arithmetic the last sample() call actually ran ('tc' may have switched to 'tc32')."""
        return getattr(self, "_ran_precision", None)

    @property
    def step_size(self):
        return self.simulate.step_size

    @step_size.setter
    def step_size(self, step_size):
        self.simulate.step_size = step_size

    @property
    def steps(self):
        return self.simulate.steps

    @steps.setter
    def steps(self, steps):
        self.simulate.steps = steps

    def _run_ensemble(self, x0, sample_size, thin, store_chain, first_chain, replay):
        from ..surrogate.extreme_learning import SurrogateGradient
        tgt = self.target_density
        from ..core.density import as_builtin
        user_target = as_builtin(tgt) is None or (
            not type(tgt).__module__.startswith("hepmc_b200.") and
            type(tgt).pot_gradient is not BuiltinDensity.pot_gradient)     # class-level override of the gradient
        if not isinstance(self.p_dist, Gaussian):
            raise TypeError("the momentum distribution must be densities.Gaussian")
        if np.any(self.p_dist.mean != 0):
            raise TypeError("the momentum distribution must have zero mean")
        mass = self.p_dist._diag()
        custom_sim = type(self.simulate) is not HamiltonLeapfrog
        if custom_sim:
            # a user-supplied simulator cannot run inside the chain kernel: per-step device path, the
            # simulator receives and returns (C, ndim) CUDA tensors (replay tapes are not available here)
            if replay is not None:
                raise TypeError("replay tapes drive the fused kernel; they are not available with a custom simulator")
            if not callable(self.simulate):
                raise TypeError("simulate must be callable as simulate(q, p) -> (q_next, p_next)")
            return ensemble.run_hmc_generic(tgt.pdf, None, x0, mass, None, None, sample_size, thin, store_chain,
                                            first_chain, simulate=self.simulate)
        if user_target:
            # user-defined target: per-step launches, the user's pdf / pot_gradient run on (C, ndim) CUDA tensors
            if replay is not None or not (callable(getattr(tgt, "pdf", None)) and callable(getattr(tgt, "pot_gradient", None))):
                raise TypeError("a user-defined HMC target needs callable pdf and pot_gradient taking (C, ndim) CUDA "
                                "tensors; replay tapes are available for the built-in densities only")
            return ensemble.run_hmc_generic(tgt.pdf, tgt.pot_gradient, x0, mass, float(self.step_size), int(self.steps),
                                            sample_size, thin, store_chain, first_chain)
        grad = vars(tgt).get("pot_gradient", None)
        counted = grad if hasattr(grad, "count") and hasattr(grad, "fn") else None
        grad = getattr(grad, "fn", grad)            # unwrap util.count_calls (counted below)
        if getattr(grad, "__self__", None) is tgt and getattr(grad, "__func__", None) is type(tgt).pot_gradient:
            grad = None                             # the density's own gradient behind the counter
        elm_blk, keep = None, None
        if grad is not None:
            if not isinstance(grad, SurrogateGradient):
                raise TypeError("a replaced pot_gradient must be a surrogate.SurrogateGradient "
                                "(ELM surrogate); arbitrary Python gradients cannot run in the kernel")
            elm_blk, keep = grad.block()
        prec = {"f64": 0, "f32": 1, "tc": 2, "tc32": 3}[self.precision]
        if elm_blk is None:
            prec = 0
        if prec == 2:
            # TF32 operands in the contraction cannot resolve a heavily cancelling weight vector:
            # measure the cancellation once and fall back to the FP32 contraction ("tc32") above the limit
            # the decision belongs to ONE surrogate (its weights) and to this precision setting: a re-trained or
            # re-attached surrogate is measured again
            key = (id(grad), id(grad.out_weights))
            if getattr(self, "_tc_key", None) != key:
                self._tc_key = key
                kappa = grad.cancellation(x0[:256])
                self._tc_mode = 3 if kappa > TC_CANCELLATION_LIMIT else 2
                if self._tc_mode == 3:
                    import warnings
                    warnings.warn("precision='tc': the surrogate's gradient sum cancels by a factor %.0f (> %.0f), which "
                                  "TF32 operands would lose; using precision='tc32' (tensor cores for the exponents, "
                                  "FP32 contraction) instead.  extreme_learning_train(..., ridge=1e-6) avoids the "
                                  "cancellation." % (kappa, TC_CANCELLATION_LIMIT), RuntimeWarning, stacklevel=3)
            prec = self._tc_mode
        self._ran_precision = ("f64", "f32", "tc", "tc32")[prec]
        out = ensemble.run_hmc(tgt._block(), elm_blk, prec, x0, mass, self.step_size, self.steps,
                               sample_size, thin, store_chain, first_chain, None, replay,
                               getattr(self, "_total_chains", None))
        if counted is not None:
            # every transition evaluates the gradient steps + 1 times (simulation.py:34-42), one row per
            # call in the reference: count += chains * transitions * (steps + 1)
            counted.count += x0.shape[0] * (sample_size - 1) * (int(self.steps) + 1)
        del keep
        return out

    def sample(self, sample_size, initial, out_mask=None, log_every=5000, **kwargs):
        sample = super().sample(sample_size, initial, out_mask, log_every, **kwargs)
        sample.target = self.target_density
        return sample
