"""Factorised CPU restatement of the CPHF hot path for sizes the reference cannot hold.

TEST INFRASTRUCTURE ONLY (like oracle/pmr_oracle.py): imported by tests/, by
``__graft_entry__.smoke()`` and by bench.py's ``check`` / cpu legs -- never by the product
package.  NumPy + SciPy/LAPACK, no CUDA.

Why it exists (SURVEY.md 7.3 item 1, 8(c) last row): at nbasis = 512 the reference needs the
550 GB AO tensor, the 26 GB explicit super-Hessian and its inverse per frequency, so it cannot
be run to check BASELINE configs 3/4/5.  The synthetic AO integrals are
``(mu nu|la si) = sum_P B[P,mu,nu] B[P,la,si]`` (pymolresponse_b200/synthetic.py), so the same
numbers the reference would compute are reachable without any n^4 object:

* MO integrals: ``(pq|rs) = sum_P (C_p^T B_P C_q)(C_r^T B_P C_s)`` -- what
  ``AO2MO.transform`` (ao2mo.py:35-50) gives for that I, quarter by quarter.
* A / B_ref blocks: the arithmetic of explicit_equations_partial.py:8-214 in the same order
  (``2*x``, ``-``, ``+ diag``), formed per occupied row block so nothing larger than the
  result lives in memory.
* Response: the super-Hessian equations of solvers.py:179-231 + 401-425,
  ``[[A - w, B],[B, A + w]] [X;Y] = [g; s g]``, reduced with K = A + B_ref, M = A - B_ref,
  p = X + Y, m = X - Y to  ``K p - w m = (1+s) g``, ``M m - w p = (1-s) g``  (SURVEY.md A.3;
  UHF: alpha and beta stacked, A.4), solved with LAPACK Cholesky at size nov.
* Results: ``results = pref * sum_sigma P R^T`` (cphf.py:165-253, utils.py:17-27), pref = 2
  (RHF) / 1 (UHF).

It is deliberately a DIFFERENT formulation from the CUDA path's device engine only in the
parts that do not matter for the check (LAPACK instead of the hand-written kernels, explicit
K^-1 by ``cho_solve`` on identity columns), and it is pinned before use:
tests/test_oracle_factorised.py proves it equal to ``pmr_oracle.run_cphf`` (the pinned
restatement of the reference's explicit-inverse algorithm) and, in the build container, to the
unmodified reference itself (tests/test_golden_regeneration.py) for RHF and UHF, real and
imaginary operators, static and dynamic frequencies (blocks <= 1e-12 * max, results <= 1e-10).

All big products run in row chunks: one 10.9 GB DGEMM segfaulted inside OpenBLAS in this
image, and the blocked Cholesky below keeps every LAPACK/BLAS call small
for the same reason.
"""

from __future__ import annotations

import numpy as np
import scipy.linalg

CHUNK = 4096  # rows per BLAS call


# --------------------------------------------------------------------------- MO integrals
def mo_factor(B, Cl, Cr):
    """F[P, l, r] = sum_{mu,nu} Cl[mu,l] B[P,mu,nu] Cr[nu,r]  -> (naux, dl, dr)."""
    naux, n, _ = B.shape
    half = np.empty((naux, n, Cr.shape[1]))
    for p0 in range(0, naux, 64):
        half[p0:p0 + 64] = B[p0:p0 + 64] @ Cr
    return np.einsum("ml,pmr->plr", Cl, half, optimize=True)


def _gram(FL, FR):
    """out[x, y] = sum_P FL[P, x] FR[P, y], in row chunks."""
    nx, ny = FL.shape[1], FR.shape[1]
    out = np.empty((nx, ny))
    FLt = np.ascontiguousarray(FL.T)
    for r0 in range(0, nx, CHUNK):
        r1 = min(nx, r0 + CHUNK)
        np.dot(FLt[r0:r1], FR, out=out[r0:r1])
    return out


def mo_blocks_rhf(B, C, nocc):
    """(ovov[o,v,o,v], oovv[o,o,v,v]) as AO2MO.perform_rhf_partial (ao2mo.py:58-70) returns them."""
    n = B.shape[1]
    o, v = nocc, n - nocc
    Co, Cv = C[:, :o], C[:, o:]
    naux = B.shape[0]
    Fov = mo_factor(B, Co, Cv).reshape(naux, o * v)
    Foo = mo_factor(B, Co, Co).reshape(naux, o * o)
    Fvv = mo_factor(B, Cv, Cv).reshape(naux, v * v)
    ovov = _gram(Fov, Fov).reshape(o, v, o, v)
    oovv = _gram(Foo, Fvv).reshape(o, o, v, v)
    return ovov, oovv


def mo_blocks_uhf(B, C, occupations):
    """The 6-tuple of interfaces/pyscf/ao2mo.py:71-109 (ovov aaaa, aabb, bbaa, bbbb; oovv aaaa, bbbb)."""
    na, va, nb, vb = occupations
    naux = B.shape[0]
    Ca, Cb = C[0], C[1]
    Fov_a = mo_factor(B, Ca[:, :na], Ca[:, na:]).reshape(naux, na * va)
    Fov_b = mo_factor(B, Cb[:, :nb], Cb[:, nb:]).reshape(naux, nb * vb)
    Foo_a = mo_factor(B, Ca[:, :na], Ca[:, :na]).reshape(naux, na * na)
    Foo_b = mo_factor(B, Cb[:, :nb], Cb[:, :nb]).reshape(naux, nb * nb)
    Fvv_a = mo_factor(B, Ca[:, na:], Ca[:, na:]).reshape(naux, va * va)
    Fvv_b = mo_factor(B, Cb[:, nb:], Cb[:, nb:]).reshape(naux, vb * vb)
    return (
        _gram(Fov_a, Fov_a).reshape(na, va, na, va),
        _gram(Fov_a, Fov_b).reshape(na, va, nb, vb),
        _gram(Fov_b, Fov_a).reshape(nb, vb, na, va),
        _gram(Fov_b, Fov_b).reshape(nb, vb, nb, vb),
        _gram(Foo_a, Fvv_a).reshape(na, na, va, va),
        _gram(Foo_b, Fvv_b).reshape(nb, nb, vb, vb),
    )


# --------------------------------------------------------------------------- M and K
def _ediff(E_MO, nocc):
    e = np.diag(E_MO)
    return (e[None, nocc:] - e[:nocc, None]).reshape(-1)   # utils.py:326-347, index i*v + a


def _fill_mk_same_spin(Mout, Kout, r0, c0, ovov, oovv, ediff, c_ovov, singlet=True, tda=False):
    """Same-spin / closed-shell block, one occupied row i at a time.

    RHF singlet (c_ovov = 2; explicit_equations_partial.py:34-38, 92-96):
        A = 2 ovov - oovv^{ijab->iajb} + diag,  B_ref = -(2 ovov - ovov^{swap(1,3)})
    UHF same spin (c_ovov = 1; :145-149, :190-194):
        A = ovov - oovv^T + diag,               B_ref = -(ovov - ovov^{swap(1,3)})
    triplet (singlet=False; :66-70, :114-118):
        A = -oovv^T + diag,                     B_ref = +ovov^{swap(1,3)}
    TDA: B_ref = 0 (solvers.py:208,225).  M = A - B_ref, K = A + B_ref.
    """
    o, v = ovov.shape[0], ovov.shape[1]
    for i in range(o):
        rows = slice(r0 + i * v, r0 + (i + 1) * v)
        direct = ovov[i].reshape(v, o * v)                                  # [a, (j b)]
        swapped = ovov[i].transpose(2, 1, 0).reshape(v, o * v)              # (ib|ja) -> [a, (j b)]
        exch = oovv[i].transpose(1, 0, 2).reshape(v, o * v)                 # (ij|ab) -> [a, (j b)]
        if singlet:
            A = c_ovov * direct - exch
            Bm = -(c_ovov * direct - swapped)
        else:
            A = -exch
            Bm = swapped.copy()
        if tda:
            Bm = np.zeros_like(A)
        idx = np.arange(v)
        A[idx, i * v + idx] += ediff[i * v:(i + 1) * v]
        Mout[rows, c0:c0 + o * v] = A - Bm
        Kout[rows, c0:c0 + o * v] = A + Bm


def mk_rhf(E_MO, ovov, oovv, hamiltonian="rpa", spin="singlet"):
    """M = A - B_ref, K = A + B_ref for an RHF reference (solvers.py:189-228); TDA: B_ref = 0."""
    o, v = ovov.shape[0], ovov.shape[1]
    nov = o * v
    M = np.empty((nov, nov))
    K = np.empty((nov, nov))
    _fill_mk_same_spin(M, K, 0, 0, ovov, oovv, _ediff(E_MO, o), 2.0, singlet=(spin == "singlet"),
                       tda=(hamiltonian == "tda"))
    return M, K


def mk_uhf(E, tei6, occupations):
    """Joint alpha+beta M and K from the blocks of solvers.py:233-370 (RPA singlet):
    same-spin diagonal blocks, opposite-spin A_os = ovov_xxyy, B_os = -ovov_xxyy."""
    na, va, nb, vb = occupations
    nova, novb = na * va, nb * vb
    N = nova + novb
    ovov_aaaa, ovov_aabb, ovov_bbaa, ovov_bbbb, oovv_aaaa, oovv_bbbb = tei6
    M = np.empty((N, N))
    K = np.empty((N, N))
    _fill_mk_same_spin(M, K, 0, 0, ovov_aaaa, oovv_aaaa, _ediff(E[0], na), 1.0)
    _fill_mk_same_spin(M, K, nova, nova, ovov_bbbb, oovv_bbbb, _ediff(E[1], nb), 1.0)
    ab = ovov_aabb.reshape(nova, novb)
    ba = ovov_bbaa.reshape(novb, nova)
    M[:nova, nova:] = ab - (-ab)        # A_os - B_os
    K[:nova, nova:] = ab + (-ab)
    M[nova:, :nova] = ba - (-ba)
    K[nova:, :nova] = ba + (-ba)
    return M, K


# --------------------------------------------------------------------------- dense SPD algebra
def _small_tri_inverse(Ljj):
    return scipy.linalg.solve_triangular(Ljj, np.eye(Ljj.shape[0]), lower=True, check_finite=False)


def cholesky_blocked(A, nb=1024):
    """In-place lower Cholesky (right-looking, LAPACK dpotrf on the diagonal blocks, GEMMs elsewhere);
    returns A, whose lower triangle holds L.  Raises LinAlgError if A is not positive definite."""
    N = A.shape[0]
    for j in range(0, N, nb):
        je = min(N, j + nb)
        Ljj = scipy.linalg.cholesky(A[j:je, j:je], lower=True, check_finite=False)
        A[j:je, j:je] = Ljj
        if je == N:
            break
        WT = _small_tri_inverse(Ljj).T.copy()
        for r0 in range(je, N, CHUNK):
            r1 = min(N, r0 + CHUNK)
            A[r0:r1, j:je] = A[r0:r1, j:je] @ WT
        for r0 in range(je, N, CHUNK):
            r1 = min(N, r0 + CHUNK)
            # lower trapezoid of the trailing matrix only
            A[r0:r1, je:r1] -= A[r0:r1, j:je] @ A[je:r1, j:je].T
    return A


class Factor:
    """Lower Cholesky factor L (N, N) plus the explicit inverses of its nb x nb diagonal blocks, so
    that every substitution step is a GEMM (scipy's solve_triangular runs at a few GFLOP/s here)."""

    def __init__(self, L, nb=1024):
        self.L, self.nb, self.N = L, nb, L.shape[0]
        self.dinv = {j: _small_tri_inverse(L[j:min(self.N, j + nb), j:min(self.N, j + nb)])
                     for j in range(0, self.N, nb)}

    def forward(self, Bm, start=0):
        """Bm <- L[start:, start:]^-1 Bm (Bm has N - start rows; start is a multiple of nb)."""
        L, nb, N = self.L, self.nb, self.N
        for j in range(start, N, nb):
            je = min(N, j + nb)
            Bm[j - start:je - start] = self.dinv[j] @ Bm[j - start:je - start]
            for r0 in range(je, N, CHUNK):
                r1 = min(N, r0 + CHUNK)
                Bm[r0 - start:r1 - start] -= L[r0:r1, j:je] @ Bm[j - start:je - start]
        return Bm

    def backward(self, Bm):
        """Bm <- L^-T Bm."""
        L, nb, N = self.L, self.nb, self.N
        for j in reversed(range(0, N, nb)):
            je = min(N, j + nb)
            Bm[j:je] = self.dinv[j].T @ Bm[j:je]
            for r0 in range(0, j, CHUNK):
                r1 = min(j, r0 + CHUNK)
                Bm[r0:r1] -= L[j:je, r0:r1].T @ Bm[j:je]
        return Bm

    def solve(self, Bm):
        """(L L^T)^-1 B as a new (N, k) array."""
        Bm = np.array(Bm, dtype=np.float64, order="C", copy=True)
        return self.backward(self.forward(Bm))

    def tri_inverse(self):
        """W = L^-1 (lower triangular) by forward substitution on identity block columns, skipping
        the rows above each block that stay zero (N^3/3 flops)."""
        N, nb = self.N, self.nb
        W = np.zeros((N, N))
        for c0 in range(0, N, nb):
            c1 = min(N, c0 + nb)
            E = np.zeros((N - c0, c1 - c0))
            E[np.arange(c1 - c0), np.arange(c1 - c0)] = 1.0
            W[c0:, c0:c1] = self.forward(E, start=c0)
        return W

    def inverse_lower(self):
        """(L L^T)^-1 = W^T W with W = L^-1; LOWER triangle filled block-wise (blocks strictly above
        the diagonal stay zero): its consumer S(w) = M - w^2 K^-1 is only read by the lower
        Cholesky."""
        N, nb = self.N, self.nb
        W = self.tri_inverse()
        out = np.zeros((N, N))
        for i0 in range(0, N, nb):
            i1 = min(N, i0 + nb)
            Wi = np.ascontiguousarray(W[i0:, i0:i1].T)        # rows k >= i0 are the non-zero ones
            for j0 in range(0, i1, CHUNK):
                j1 = min(i1, j0 + CHUNK)
                out[i0:i1, j0:j1] = Wi @ W[i0:, j0:j1]
        return out


def cho_solve_blocked(L, Bm, nb=1024):
    return Factor(L, nb).solve(Bm)


def spd_inverse(L, nb=1024):
    return Factor(L, nb).inverse_lower()


# --------------------------------------------------------------------------- response
def solve_half_size(M, K, g_cols, signs, frequencies):
    """All frequencies / columns of  K p - w m = (1+s) g,  M m - w p = (1-s) g.

    g_cols: (N, nc) property gradients (first half of the supervector [g; s g], operators.py:75-101),
    signs[c] = -1 real / +1 imaginary.  Returns xy (nf, nc, 2N) with [X; Y] = [(p+m)/2; (p-m)/2]
    -- the response supervectors of solvers.py:407-425.  M and K are overwritten by their factors.
    """
    N, nc = g_cols.shape
    s = np.asarray(signs, dtype=np.float64)
    freqs = [float(w) for w in frequencies]
    any_dyn = any(w != 0.0 for w in freqs)
    need_K = any_dyn or bool(np.any(s > 0))     # static real operators only need M (p = 0)
    Mkeep = M.copy() if any_dyn else None
    FK = Factor(cholesky_blocked(K)) if need_K else None
    Kinv = FK.inverse_lower() if any_dyn else None
    FM = Factor(cholesky_blocked(M))
    Kinv_g = FK.solve(g_cols) if need_K else np.zeros_like(g_cols)
    xy = np.zeros((len(freqs), nc, 2 * N))
    for f, w in enumerate(freqs):
        rhs_m = (1.0 - s)[None, :] * g_cols + w * (1.0 + s)[None, :] * Kinv_g
        if w == 0.0:
            m = FM.solve(rhs_m)
        else:
            S = Mkeep - (w * w) * Kinv          # lower triangle is what matters
            m = Factor(cholesky_blocked(S)).solve(rhs_m)
            del S
        p = FK.solve((1.0 + s)[None, :] * g_cols + w * m) if need_K else np.zeros_like(m)
        xy[f, :, :N] = (0.5 * (p + m)).T
        xy[f, :, N:] = (0.5 * (p - m)).T
    return xy


def property_gradients(ao_integrals, C, nocc):
    """g[c, i*v + a] = (C_virt^T O_c C_occ)[a, i]  (operators.py:85-96, utils.py:68-70)."""
    Co, Cv = C[:, :nocc], C[:, nocc:]
    out = []
    for O in ao_integrals:
        mat = Cv.T @ (O @ Co)                  # (v, o)
        out.append(mat.reshape(-1, order="F"))  # index i*v + a
    return np.asarray(out)


def sample_indices(N, count, seed=0):
    """Seeded (row, col) probes of the lower triangle (what the device factorisations read) plus a
    sweep of the diagonal: the M / K elements compared between the CPU restatement and the GPU."""
    rng = np.random.default_rng(seed)
    rr = rng.integers(0, N, count)
    cc = rng.integers(0, N, count)
    lo, hi = np.maximum(rr, cc), np.minimum(rr, cc)
    diag = np.arange(0, N, max(1, N // 64))
    return np.concatenate((lo, diag)).astype(np.int64), np.concatenate((hi, diag)).astype(np.int64)


def run_cphf_factorised(B, C, E, occupations, operators, frequencies, hamiltonian="rpa",
                        spin="singlet", return_mk_samples=0, seed=0):
    """CPHF.run with an exact solver (cphf.py:56-66) from the AO factor B.

    operators: list of dicts {ao_integrals (ncomp,n,n), is_imaginary}.  Returns dict(results: list
    of (sum ncomp, sum ncomp) per frequency, xy_alph / xy_beta: (nf, ncols, 2 nov_s) supervectors,
    mk_samples: optional list of (row, col, M, K) probes for Hessian-element checks).
    """
    if C.ndim == 2:
        C = C[None]
    if E.ndim == 2:
        E = E[None]
    is_uhf = C.shape[0] == 2
    na, va, nb, vb = occupations
    if is_uhf:
        assert hamiltonian == "rpa" and spin == "singlet"
        tei = mo_blocks_uhf(B, C, occupations)
        M, K = mk_uhf(E, tei, occupations)
        del tei
    else:
        ovov, oovv = mo_blocks_rhf(B, C[0], na)
        M, K = mk_rhf(E[0], ovov, oovv, hamiltonian, spin)
        del ovov, oovv
    N = M.shape[0]
    samples = []
    if return_mk_samples:
        rr, cc = sample_indices(N, return_mk_samples, seed)
        samples = [(int(r), int(c), float(M[r, c]), float(K[r, c])) for r, c in zip(rr, cc)]
        samples.append(("max_abs", "max_abs", float(np.abs(M).max()), float(np.abs(K).max())))
    cols, signs, spans = [], [], []
    for op in operators:
        ga = property_gradients(op["ao_integrals"], C[0], na)
        if is_uhf:
            gb = property_gradients(op["ao_integrals"], C[1], nb)
            ga = np.concatenate((ga, gb), axis=1)
        spans.append((len(cols), len(cols) + ga.shape[0]))
        for row in ga:
            cols.append(row)
            signs.append(+1 if op.get("is_imaginary", False) else -1)
    G = np.asarray(cols).T.copy()                      # (N, ncols)
    xy = solve_half_size(M, K, G, signs, frequencies)  # (nf, ncols, 2N)
    del M, K
    sg = np.asarray(signs, dtype=np.float64)
    results = []
    pref = 1.0 if is_uhf else 2.0
    for f in range(len(frequencies)):
        X, Y = xy[f][:, :N], xy[f][:, N:]
        # results[a, b] = pref * [g_a; s_a g_a] . [X_b; Y_b]   (utils.py:17-27, cphf.py:249-252)
        results.append(pref * (G.T @ X.T + (sg[:, None] * G.T) @ Y.T))
    out = {"results": results, "signs": signs, "spans": spans, "mk_samples": samples}
    if is_uhf:
        nova = na * va
        Xa, Xb = xy[:, :, :nova], xy[:, :, nova:N]
        Ya, Yb = xy[:, :, N:N + nova], xy[:, :, N + nova:]
        out["xy_alph"] = np.concatenate((Xa, Ya), axis=2)
        out["xy_beta"] = np.concatenate((Xb, Yb), axis=2)
    else:
        out["xy_alph"] = xy
    return out
