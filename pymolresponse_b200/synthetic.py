"""Seeded synthetic inputs for the linear-response hot path.

NumPy only (no CUDA): shared by the tests, the oracle's golden-vector generator and
``bench.py``.  "Random-SPD AO ERIs": ``I[mu,nu,la,si] = sum_P B[P,mu,nu] B[P,la,si]`` with a
seeded, (mu,nu)-symmetrised factor ``B`` so that I has the 8-fold symmetry of real ERIs and is
PSD as an (n^2 x n^2) matrix.  MO coefficients are an orthogonal matrix, orbital energies are
gapped so that A+B and A-B are SPD (first pole > 1 a.u.).
"""

from __future__ import annotations

import numpy as np

SEED_ERI = 0
SEED_ALPHA = 1
SEED_BETA = 2
SEED_OPERATOR = 5
ERI_SCALE = 0.05


def eri_factor(n: int, naux: int | None = None, seed: int = SEED_ERI) -> np.ndarray:
    """B[P, mu, nu], symmetric in (mu, nu), scaled by ERI_SCALE/sqrt(naux)."""
    naux = n if naux is None else naux
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((naux, n, n))
    B = 0.5 * (B + B.transpose(0, 2, 1))
    B *= ERI_SCALE / np.sqrt(naux)
    return B


def eri_from_factor(B: np.ndarray, mu_lo: int = 0, mu_hi: int | None = None, chunk: int = 4096):
    """I[mu_lo:mu_hi] = B^T B built in row chunks (one huge DGEMM segfaults OpenBLAS here)."""
    naux, n, _ = B.shape
    mu_hi = n if mu_hi is None else mu_hi
    Bf = B.reshape(naux, n * n)
    rows = Bf[:, mu_lo * n : mu_hi * n]
    out = np.empty(((mu_hi - mu_lo) * n, n * n))
    for r0 in range(0, out.shape[0], chunk):
        r1 = min(r0 + chunk, out.shape[0])
        np.dot(rows[:, r0:r1].T, Bf, out=out[r0:r1])
    return out.reshape(mu_hi - mu_lo, n, n, n)


def mo_coefficients(n: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return Q


def orbital_energies(n: int, nocc: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed + 1000)
    occ = np.sort(rng.uniform(-2.0, -0.5, nocc))
    virt = np.sort(rng.uniform(0.5, 3.0, n - nocc))
    return np.concatenate((occ, virt))


def operator_integrals(n: int, ncomp: int = 3, imaginary: bool = False, seed: int = SEED_OPERATOR):
    rng = np.random.default_rng(seed + (100 if imaginary else 0))
    O = rng.standard_normal((ncomp, n, n))
    if imaginary:
        O = 0.5 * (O - O.transpose(0, 2, 1))
    else:
        O = 0.5 * (O + O.transpose(0, 2, 1))
    return O / n


def rhf_case(n: int, nocc: int, build_eri: bool = True):
    """dict(C (1,n,n), E (1,n,n) diagonal matrices, occupations, I or None, B factor)."""
    B = eri_factor(n)
    C = mo_coefficients(n, SEED_ALPHA)[None]
    E = np.diag(orbital_energies(n, nocc, SEED_ALPHA))[None]
    return {
        "C": C,
        "E": E,
        "occupations": (nocc, n - nocc, nocc, n - nocc),
        "I": eri_from_factor(B) if build_eri else None,
        "B": B,
    }


def uhf_case(n: int, nocc_a: int, nocc_b: int, build_eri: bool = True):
    B = eri_factor(n)
    C = np.stack((mo_coefficients(n, SEED_ALPHA), mo_coefficients(n, SEED_BETA)), axis=0)
    E = np.stack(
        (
            np.diag(orbital_energies(n, nocc_a, SEED_ALPHA)),
            np.diag(orbital_energies(n, nocc_b, SEED_BETA)),
        ),
        axis=0,
    )
    return {
        "C": C,
        "E": E,
        "occupations": (nocc_a, n - nocc_a, nocc_b, n - nocc_b),
        "I": eri_from_factor(B) if build_eri else None,
        "B": B,
    }


# ---- algorithmic FP64 flop counts; numerators of the roofline ----


def flops_ao2mo_rhf_partial(n: int, o: int) -> float:
    v = n - o
    return 2.0 * n**4 * o + 2 * 2.0 * n**3 * o**2 + 2 * 2.0 * n**2 * o**2 * v + 2 * 2.0 * n * o**2 * v**2


def flops_solve_rhf(nov: int, n_static: int, n_dynamic: int, nrhs: int) -> float:
    """potrf(M or K) for the static solve; potrf(K) + K^-1 (2u) + one potrf(S(w)) per dynamic w."""
    u = nov**3 / 3.0
    f = 0.0
    if n_dynamic > 0:
        f += 3.0 * u + n_dynamic * u
    if n_static > 0:
        f += u
    f += 4.0 * nov**2 * nrhs * (n_static + n_dynamic)
    return f


def bytes_assemble(nov: int) -> float:
    return 4.0 * nov * nov * 8.0
