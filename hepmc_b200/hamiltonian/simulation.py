"""Leapfrog integrator settings (mirror of hepmc/hamiltonian/simulation.py:4-46).

In the ensemble kernels the leapfrog loop runs in registers with the reference's
structure: g = dU(q); L x { p -= eps/2 g; q += eps dK(p); g = dU(q); p -= eps/2 g }
(half steps not merged, L+1 gradient evaluations).  This object carries step size /
step count for HamiltonianUpdate.
"""

class HamiltonLeapfrog(object):

    def __init__(self, pot_gradient, kin_gradient, step_size, steps):
        self.kin_gradient = kin_gradient
        self.pot_gradient = pot_gradient
        self.step_size = step_size
        self.steps = steps

    def __call__(self, q_init, p_init):
        raise NotImplementedError(
            "hepmc_b200 integrates trajectories inside the ensemble kernel (hepmc_hmc_chains); "
            "use HamiltonianUpdate.sample -- there is no host-side leapfrog (no CPU fallback).")
