"""Phase-space generators as Distributions: `Rambo(nparticles, E_CM)` and
`RamboOnDiet(nparticles, E_CM)` (reference: hepmc/core/densities/rambo.py:6-63).

`rvs(n)` draws this rank's chunk-aligned share of n phase-space points straight from the
mapping kernel (uniforms never leave the registers; point i uses Philox counter i, so the union
over ranks does not depend on the number of GPUs); `pdf` is the constant 1/Phi_n.
"""
from ..density import Distribution
from .. import phase_space
from ... import _lib, config, parallel


class _PhaseSpaceDistribution(Distribution):
    """Shared body: only the mapping class differs between the two generators."""

    mapping_class = None

    def __init__(self, nparticles, E_CM):
        self.mapping = self.mapping_class(E_CM, nparticles)
        Distribution.__init__(self, self.mapping.ndim, False)

    def rvs(self, sample_size):
        total = int(sample_size)
        first, count, _ = parallel.local_span(total, _lib.default_chunk(total))
        return config.deliver(self.mapping.generate(count, first_index=first))

    def pdf(self, xs):
        return self.mapping.pdf(xs)

    # the two knobs of the mapping, exposed like the reference does
    e_cm = property(lambda self: self.mapping.e_cm,
                    lambda self, value: setattr(self.mapping, "e_cm", value))
    nparticles = property(lambda self: self.mapping.nparticles,
                          lambda self, value: setattr(self, "mapping", self.mapping_class(self.e_cm, value)))


class Rambo(_PhaseSpaceDistribution):
    mapping_class = phase_space.Rambo


class RamboOnDiet(_PhaseSpaceDistribution):
    mapping_class = phase_space.RamboOnDiet
