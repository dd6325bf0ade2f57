"""CPU: the factorised restatement (oracle/pmr_factorised.py) is pinned before it is trusted at the
benchmark sizes.  It must equal the pinned NumPy oracle (explicit super-Hessian + inverse per
frequency, the reference's algorithm) on the same seeded inputs: MO blocks and Hessian elements
<= 1e-12 * max, property tensors and response vectors <= 1e-10.  The live-reference comparison at
n = 128 / 192 is in tests/test_golden_regeneration.py (build container only)."""

import numpy as np
import pytest

from oracle import pmr_factorised as fac
from oracle import pmr_oracle as orc
from pymolresponse_b200 import synthetic


def _ops(n):
    return [{"ao_integrals": synthetic.operator_integrals(n, 3, False), "is_imaginary": False},
            {"ao_integrals": synthetic.operator_integrals(n, 3, True), "is_imaginary": True}]


@pytest.mark.parametrize("n,o", [(24, 3), (56, 7)])
def test_rhf_blocks_and_mk(n, o):
    case = synthetic.rhf_case(n, o)
    ovov, oovv = fac.mo_blocks_rhf(case["B"], case["C"][0], o)
    r_ovov, r_oovv = orc.ao2mo_rhf_partial(case["I"], case["C"], o)
    assert np.abs(ovov - r_ovov).max() <= 1e-12 * np.abs(r_ovov).max()
    assert np.abs(oovv - r_oovv).max() <= 1e-12 * np.abs(r_oovv).max()
    for ham in ("rpa", "tda"):
        for spin in ("singlet", "triplet"):
            A, B = orc.rhf_ab(case["E"], (r_ovov, r_oovv), "partial", o, ham, spin)
            M, K = fac.mk_rhf(case["E"][0], r_ovov, r_oovv, ham, spin)
            # same operands, same operation order: identical bits
            assert np.array_equal(M, A - B), (ham, spin)
            assert np.array_equal(K, A + B), (ham, spin)


@pytest.mark.parametrize("ham,spin", [("rpa", "singlet"), ("rpa", "triplet"), ("tda", "singlet")])
def test_rhf_results_equal_explicit_inverse_oracle(ham, spin):
    n, o = 48, 6
    case = synthetic.rhf_case(n, o)
    freqs = [0.0, 0.1, 0.35]
    tei = orc.ao2mo_rhf_partial(case["I"], case["C"], o)
    ref = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial", _ops(n), freqs, ham, spin)
    got = fac.run_cphf_factorised(case["B"], case["C"], case["E"], case["occupations"], _ops(n), freqs,
                                  ham, spin, return_mk_samples=20)
    nov = o * (n - o)
    for f in range(len(freqs)):
        np.testing.assert_allclose(got["results"][f], ref["results"][f], rtol=0, atol=1e-10)
        for k, (c0, c1) in enumerate(got["spans"]):
            want = ref["rspvecs_alph"][k][f][:, :, 0]
            np.testing.assert_allclose(got["xy_alph"][f, c0:c1], want, rtol=0, atol=1e-10)
    A, B = orc.rhf_ab(case["E"], tei, "partial", o, ham, spin)
    for r, c, mv, kv in got["mk_samples"][:-1]:
        assert abs(mv - (A - B)[r, c]) <= 1e-12 * np.abs(A).max()
        assert abs(kv - (A + B)[r, c]) <= 1e-12 * np.abs(A).max()
    assert got["xy_alph"].shape == (3, 6, 2 * nov)


def test_uhf_results_equal_explicit_inverse_oracle():
    n, na, nb = 30, 7, 5
    case = synthetic.uhf_case(n, na, nb)
    occ = case["occupations"]
    freqs = [0.0, 0.09]
    tei = orc.ao2mo_uhf_partial(case["I"], case["C"], occ)
    mine = fac.mo_blocks_uhf(case["B"], case["C"], occ)
    for got, want in zip(mine, tei):
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    ref = orc.run_cphf(case["C"], case["E"], occ, tei, "partial", _ops(n), freqs)
    got = fac.run_cphf_factorised(case["B"], case["C"], case["E"], occ, _ops(n), freqs)
    for f in range(len(freqs)):
        np.testing.assert_allclose(got["results"][f], ref["results"][f], rtol=0, atol=1e-10)
        for k, (c0, c1) in enumerate(got["spans"]):
            np.testing.assert_allclose(got["xy_alph"][f, c0:c1], ref["rspvecs_alph"][k][f][:, :, 0],
                                       rtol=0, atol=1e-10)
            np.testing.assert_allclose(got["xy_beta"][f, c0:c1], ref["rspvecs_beta"][k][f][:, :, 0],
                                       rtol=0, atol=1e-10)


def test_blocked_spd_algebra_against_lapack():
    rng = np.random.default_rng(11)
    N = 2500                       # > one 1024 panel and > one 2048 row chunk
    X = rng.standard_normal((N, N + 50))
    A = X @ X.T / N + np.eye(N)
    L = fac.cholesky_blocked(A.copy(), nb=1024)
    Lref = np.linalg.cholesky(A)
    assert np.abs(np.tril(L) - Lref).max() <= 1e-12
    Bm = rng.standard_normal((N, 5))
    np.testing.assert_allclose(fac.cho_solve_blocked(L, Bm), np.linalg.solve(A, Bm), rtol=0, atol=1e-11)
    Ainv = fac.spd_inverse(L, nb=700)
    assert np.abs(np.triu(Ainv, 701)).max() == 0.0          # only the lower triangle is filled
    Afull = np.tril(Ainv) + np.tril(Ainv, -1).T
    assert np.abs(Afull @ A - np.eye(N)).max() <= 1e-11
    bad = A.copy()
    bad[10, 10] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        fac.cholesky_blocked(bad)
