"""Dense exact diagonalisation for the reference's ED-comparison jobs, redone in numpy.

Restates /root/reference/test/ed/ed.jl:6-157 and test/ed/magnet.jl:1-116 (Hamiltonian, thermal ensemble,
Energy, SpecificHeat, Mag, AbsMag, Mag2, Mag4, BinderRatio and the Kubo-integral MagChi).  Test infrastructure."""
import numpy as np

from sse_b200.operators import spin_operators


def _lift(dims, pos, op):
    left = int(np.prod(dims[:pos]))
    right = int(np.prod(dims[pos + 1:]))
    return np.kron(np.kron(np.eye(left), op), np.eye(right))


def _spin(dims, pos, idx):
    splus, sz = spin_operators(dims[pos])
    if idx == 1:
        return _lift(dims, pos, 0.5 * (splus + splus.T)).astype(complex)
    if idx == 2:
        return _lift(dims, pos, 0.5j * (splus - splus.T))
    return _lift(dims, pos, sz).astype(complex)


def hamiltonian(magnet):
    """test/ed/magnet.jl:10-26"""
    dims = [s.spin_states for s in magnet.site_params]
    D = int(np.prod(dims))
    H = np.zeros((D, D), dtype=complex)
    for bond, p in zip(magnet.lattice.bonds, magnet.bond_params):
        i, j = bond.i - 1, bond.j - 1
        heis = sum(_spin(dims, i, a) @ _spin(dims, j, a) for a in (1, 2, 3))
        szi, szj = _spin(dims, i, 3), _spin(dims, j, 3)
        sxi, sxj = _spin(dims, i, 1), _spin(dims, j, 1)
        H += (p.J * heis + p.J * p.d * szi @ szj + p.hz[0] * szi + p.hz[1] * szj + p.Dx[0] * sxi @ sxi
              + p.Dx[1] * sxj @ sxj + p.Dz[0] * szi @ szi + p.Dz[1] * szj @ szj)
    assert np.abs(H.imag).max() < 1e-12
    return H.real


def run_ed(model, Ts, estimators=()):
    """test/ed/ed.jl:121-157 -> dict name -> array over Ts."""
    magnet = getattr(model, "inner_model", model)
    Ts = np.asarray(Ts, dtype=np.float64)
    H = hamiltonian(magnet)
    Es, psi = np.linalg.eigh(H)
    rho = np.exp(-(Es[:, None] - Es.min()) / Ts[None, :])
    rho /= rho.sum(axis=0, keepdims=True)
    N = model.normalization_site_count()
    obs = {}
    obs["Energy"] = (Es[:, None] * rho).sum(0) / N
    obs["SpecificHeat"] = ((Es[:, None] ** 2 * rho).sum(0) - (Es[:, None] * rho).sum(0) ** 2) / (Ts ** 2 * N)
    dims = [s.spin_states for s in magnet.site_params]
    for est in estimators:
        sym, _ = est.obs_symbols()
        diag = np.zeros(H.shape[0])
        for i in range(len(dims)):
            vals = np.array([model.magnetization_state(est.tag, i + 1, s) for s in range(1, dims[i] + 1)])
            diag += model.staggered_sign(est.ordering_vector, est.stagger_uc, i + 1) * np.diag(_lift(dims, i, np.diag(vals)))
        diag /= N
        Mnm = psi.T @ (diag[:, None] * psi)

        def mean_diag(d):
            w = np.einsum("in,i,in->n", psi, d, psi)
            return (w[:, None] * rho).sum(0)

        obs[sym["mag"]] = mean_diag(diag)
        m2, m4 = mean_diag(diag ** 2), mean_diag(diag ** 4)
        obs[sym["mag2"]], obs[sym["mag4"]] = m2, m4
        obs[sym["absmag"]] = mean_diag(np.abs(diag))
        obs[sym["binderratio"]] = m2 ** 2 / m4
        # integrated correlator (test/ed/ed.jl:83-99)
        dE = Es[None, :] - Es[:, None]  # [n, m] = E_m - E_n
        close = np.abs(dE) < 1e-6
        chi = np.zeros(len(Ts))
        A2 = Mnm * Mnm.T
        for it, T in enumerate(Ts):
            denom = np.where(close, 2 * T, dE)
            chi[it] = (2 * rho[:, it][:, None] / denom * A2).sum()
        obs[sym["magchi"]] = chi * N
    return obs
