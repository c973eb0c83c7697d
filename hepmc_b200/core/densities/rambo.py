"""Distribution wrapper of the RAMBO mapping (hepmc/core/densities/rambo.py:6-33)."""
from ..density import Distribution
from .. import phase_space
from ... import config, parallel


class Rambo(Distribution):

    def __init__(self, nparticles, E_CM):
        self.mapping = phase_space.Rambo(E_CM, nparticles)
        super().__init__(self.mapping.ndim, False)

    def rvs(self, sample_size):
        """This rank's share of `sample_size` phase-space points (all of them in a single
        process), drawn by global point index so the union over ranks never depends on the
        number of GPUs."""
        from ... import _lib
        n = int(sample_size)
        first, count, _ = parallel.local_span(n, _lib.default_chunk(n))
        return config.deliver(self.mapping.generate(count, first_index=first))

    def pdf(self, xs):
        return self.mapping.pdf(xs)

    @property
    def e_cm(self):
        return self.mapping.e_cm

    @e_cm.setter
    def e_cm(self, value):
        self.mapping.e_cm = value

    @property
    def nparticles(self):
        return self.mapping.nparticles

    @nparticles.setter
    def nparticles(self, value):
        self.mapping = phase_space.Rambo(self.e_cm, value)


class RamboOnDiet(Distribution):
    """Distribution wrapper of the RamboOnDiet mapping (densities/rambo.py:36-63)."""

    def __init__(self, nparticles, E_CM):
        self.mapping = phase_space.RamboOnDiet(E_CM, nparticles)
        super().__init__(self.mapping.ndim, False)

    def rvs(self, sample_size):
        from ... import _lib
        n = int(sample_size)
        first, count, _ = parallel.local_span(n, _lib.default_chunk(n))
        return config.deliver(self.mapping.generate(count, first_index=first))

    def pdf(self, xs):
        return self.mapping.pdf(xs)

    @property
    def e_cm(self):
        return self.mapping.e_cm

    @e_cm.setter
    def e_cm(self, value):
        self.mapping.e_cm = value

    @property
    def nparticles(self):
        return self.mapping.nparticles

    @nparticles.setter
    def nparticles(self, value):
        self.mapping = phase_space.RamboOnDiet(self.e_cm, value)
