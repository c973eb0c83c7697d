"""How well-conditioned must a trained ELM surrogate be for the tensor-core (TF32) gradient?
Acceptance rates of HMC with float64 / float32 / tc gradients and the cancellation ratio
kappa = sum_i |term_i| / |sum_i term_i| for several trainings (development script)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hepmc_b200 as H
H.set_outputs("torch")
rs = np.random.default_rng(0)
tgt0 = H.densities.UnconstrainedCamel(8)
xs = np.concatenate([1 / 3 + 0.07 * rs.standard_normal((3000, 8)), 2 / 3 + 0.07 * rs.standard_normal((3000, 8))])
vals = tgt0.pot(xs)
C, T = 20000, 6
x0 = torch.tensor(np.where(rs.random((C, 1)) < 0.5, 1 / 3, 2 / 3) + 0.05 * rs.standard_normal((C, 8)), device="cuda")
xq = x0[:512].cpu().numpy()
for nodes in (100, 1024):
    for wlo, whi in ((0.01, 1.0), (0.3, 1.0), (0.05, 0.3)):
        for ridge in (0.0, 1e-6, 1e-3):
            basis = H.surrogate.GaussianBasis(8)
            params = (rs.random((nodes, 8)), wlo + (whi - wlo) * rs.random(nodes), None)
            design = torch.as_tensor(basis.output_matrix(xs, params), device="cuda", dtype=torch.float64)
            v = torch.as_tensor(vals, device="cuda")
            bias = float(v.mean())
            if ridge == 0.0:
                w = torch.linalg.pinv(design) @ (v - bias)
            else:
                A = design.T @ design
                lam = ridge * float(torch.trace(A)) / nodes
                w = torch.linalg.solve(A + lam * torch.eye(nodes, device="cuda", dtype=torch.float64), design.T @ (v - bias))
            w = w.cpu().numpy()
            fit = float(torch.sqrt(torch.mean((design @ torch.as_tensor(w, device="cuda") + bias - v) ** 2)))
            # cancellation ratio of the gradient at typical points
            c, wd = params[0], params[1]
            diff = xq[:, None, :] - c[None, :, :]
            p = np.exp(-np.sum(diff ** 2, axis=2) / (2 * wd ** 2))
            terms = (p * (w / wd ** 2))[:, :, None] * diff
            kappa = np.median(np.sum(np.abs(terms), axis=1) / (np.abs(np.sum(terms, axis=1)) + 1e-300))
            rates = {}
            for prec in ("f64", "f32", "tc"):
                H.seed(77)
                tgt = H.densities.Camel(8)
                H.surrogate.attach_surrogate_gradient(tgt, basis, params, bias, w)
                upd = H.hamiltonian.HamiltonianUpdate(tgt, H.densities.Gaussian(8), 20, 0.01, precision=prec)
                s = upd.sample(T, x0, store_chain=False)
                rates[prec] = float(s.accepted.double().mean()) / T
            print("nodes %4d widths [%.2f,%.2f] ridge %.0e: max|w| %.1e fit rms %.2e kappa %.1e  acceptance f64 %.3f f32 %.3f tc %.3f"
                  % (nodes, wlo, whi, ridge, np.abs(w).max(), fit, kappa, rates["f64"], rates["f32"], rates["tc"]))
