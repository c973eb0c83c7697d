"""Throughput of ONE chain / small ensembles -- the reference's own calling pattern (development script).
Reference (one chain, NumPy): RWM banana-2D 6.2e3 steps/s; HMC+ELM100 770, ELM1024 311 steps/s."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H

def rate(upd, T, x0):
    upd.sample(min(T, 50), x0); torch.cuda.synchronize()
    t0 = time.perf_counter(); s = upd.sample(T, x0); torch.cuda.synchronize()
    return (T - 1) / (time.perf_counter() - t0)

chains = [int(a) for a in sys.argv[1:]] or [1, 32, 1024]
rs = np.random.default_rng(0)
for C in chains:
    met = H.DefaultMetropolis(2, H.densities.Banana(2), H.proposals.Gaussian(2, cov=10 * np.eye(2)))
    x0 = np.array([0., 10.]) if C == 1 else np.tile([0., 10.], (C, 1))
    print("C %5d RWM banana-2D                : %.3e steps/s per chain" % (C, rate(met, 20001, x0)))
    tgt = H.densities.Camel(8)
    x8 = np.full(8, 1 / 3) if C == 1 else np.full((C, 8), 1 / 3)
    hmc = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8), 20, 0.01)
    print("C %5d HMC camel-8D exact gradient   : %.3e steps/s per chain" % (C, rate(hmc, 2001, x8)))
    for nodes in (100, 1024):
        basis = H.surrogate.GaussianBasis(8)
        params = (rs.random((nodes, 8)), 0.01 + 0.99 * rs.random(nodes), None)
        w = 0.01 * rs.standard_normal(nodes)
        for prec in ("f64", "f32", "tc32", "tc"):
            tgt = H.densities.Camel(8)
            H.surrogate.attach_surrogate_gradient(tgt, basis, params, 0.0, w)
            upd = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8), 20, 0.01, precision=prec)
            T = 201 if (nodes == 1024 and prec in ("f64", "f32")) else 1001
            print("C %5d HMC camel-8D ELM %4d %-5s    : %.3e steps/s per chain" % (C, nodes, prec, rate(upd, T, x8)))

# the reference's recipe (interfaces/mc3_hmc_el.py): mixing of IS-Metropolis and surrogate HMC, one chain
for el_size in (100, 1024):
    H.seed(1)
    target = H.densities.Camel(8)
    sampler, initial = H.interfaces.mc3_hmc_el.get_sampler(target, 8, 0.4, [1 / 3, 2 / 3], [0.1, 0.1], 0.5, nintegral=4000,
                                                           mass=1., steps=20, step_size=0.01, el_size=el_size)
    print("C     1 mc3_hmc_el recipe, el_size %4d  : %.3e steps/s per chain" % (el_size, rate(sampler, 2001, initial)))
