"""Synthetic inputs of the reference's shapes: random-start ions and random S-manifold
wavefunctions as init() builds them (SU:299-337), with numpy's PCG64 instead of drand48. Host-side data
generation only -- nothing here is on the hot path."""
import numpy as np


def random_positions(n, L, seed=12345):
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(rng.uniform(0.0, L, size=(3, n)))


def random_s_state(n, n_states=12, seed=12345):
    """psi[n][S][2]: psi_0 = sqrt(r1), psi_1 = s2 sqrt(1-r1) sqrt(r2) + i s sqrt(1-r1) sqrt(1-r2)  (SU:317-332)."""
    rng = np.random.default_rng(seed + 1)
    r1, r2 = rng.uniform(size=n), rng.uniform(size=n)
    s = np.where(rng.uniform(size=n) < 0.5, -1.0, 1.0)
    s2 = np.where(rng.uniform(size=n) < 0.5, -1.0, 1.0)
    psi = np.zeros((n, n_states, 2))
    psi[:, 0, 0] = np.sqrt(r1)
    psi[:, 1, 0] = s2 * np.sqrt(1 - r1) * np.sqrt(r2)
    psi[:, 1, 1] = s * np.sqrt(1 - r1) * np.sqrt(1 - r2)
    return psi


def random_full_state(n, n_states=12, seed=12345):
    """Normalised psi with weight on every level (for kernel parity tests)."""
    rng = np.random.default_rng(seed + 2)
    psi = rng.normal(size=(n, n_states, 2))
    psi /= np.sqrt((psi ** 2).sum(axis=(1, 2)))[:, None, None]
    return psi


def maxwellian(n, sigma, seed=12345):
    rng = np.random.default_rng(seed + 3)
    return np.ascontiguousarray(rng.normal(0.0, sigma, size=(3, n)))
