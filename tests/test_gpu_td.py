"""GPU: excitation energies -- the Jacobi SVD kernel, the RPA / TDA
eigen-solvers, the TDHF / TDA drivers and the ECD wrapper against the golden vectors written by the
unmodified reference (tests/golden/td_golden.npz) and against the oracle."""

import numpy as np
import pytest

from oracle import pmr_oracle as orc

pytestmark = pytest.mark.gpu


def _nv():
    from pymolresponse_b200 import _native as nv

    return nv


@pytest.mark.parametrize("n", [1, 2, 3, 8, 33, 100, 257])
def test_jacobi_svd_kernel(n):
    """G V = U Sigma: singular values against LAPACK (relative 1e-13), V orthogonal, columns of the
    rotated G mutually orthogonal."""
    from pymolresponse_b200 import solvers

    nv = _nv()
    rng = np.random.default_rng(100 + n)
    G = rng.standard_normal((n, n)) @ np.diag(np.logspace(0, -3, n)) @ rng.standard_normal((n, n))
    Gt = nv.to_dev(np.ascontiguousarray(G.T)).clone()
    sigma, gv, Vt, sweeps = solvers._jacobi_svd(Gt)
    s = np.sort(nv.to_host(sigma))[::-1]
    want = np.linalg.svd(G, compute_uv=False)
    np.testing.assert_allclose(s, want, rtol=1e-12, atol=1e-14 * want[0])
    V = nv.to_host(Vt).T
    np.testing.assert_allclose(V.T @ V, np.eye(n), rtol=0, atol=1e-13)
    US = nv.to_host(Gt).T                       # columns sigma_p u_p
    np.testing.assert_allclose(G @ V, US, rtol=0, atol=1e-12 * want[0])
    gram = US.T @ US
    off = gram - np.diag(np.diag(gram))
    assert np.abs(off).max() <= 1e-13 * want[0] ** 2
    assert 1 <= sweeps <= 30 or n == 1


def test_jacobi_svd_symmetric_signs():
    """For a symmetric indefinite matrix g_p . v_p is the signed eigenvalue."""
    from pymolresponse_b200 import solvers

    nv = _nv()
    rng = np.random.default_rng(7)
    Q = np.linalg.qr(rng.standard_normal((40, 40)))[0]
    lam = np.concatenate((np.linspace(-2.0, -0.3, 7), np.linspace(0.1, 5.0, 33)))
    A = (Q * lam) @ Q.T
    A = 0.5 * (A + A.T)
    _sigma, gv, _Vt, _ = solvers._jacobi_svd(nv.to_dev(A).clone())
    np.testing.assert_allclose(np.sort(nv.to_host(gv)), np.sort(lam), rtol=0, atol=1e-12)


def _run_td(sysd, tda_solver, ham, spin, ecd=False):
    from pymolresponse_b200 import integrals, operators, solvers, td
    from pymolresponse_b200.core import AO2MOTransformationType, Hamiltonian, Program, Spin
    from pymolresponse_b200.properties.ecd import ECD

    cls = solvers.ExactDiagonalizationSolverTDA if tda_solver else solvers.ExactDiagonalizationSolver
    solver = cls(sysd["C"], sysd["E"], sysd["occ"])
    solver.tei_mo = sysd["tei"]
    solver.tei_mo_type = AO2MOTransformationType[sysd["ttype"]]
    driver = (td.TDA if tda_solver else td.TDHF)(solver)
    hamiltonian = {"rpa": Hamiltonian.RPA, "tda": Hamiltonian.TDA}[ham]
    spin_e = {"singlet": Spin.singlet, "triplet": Spin.triplet}[spin]
    if ecd:
        gen = integrals.IntegralsArrays({"angmom_common_gauge": sysd["Oimag"], "dipole": sysd["Oreal"]})
        prop = ECD(Program.Arrays, gen, driver)
        prop.form_operators()
        prop.run(hamiltonian=hamiltonian, spin=spin_e)
        prop.form_results()
        return solver, driver, prop
    driver.add_operator(operators.Operator(label="angmom", is_imaginary=True,
                                           ao_integrals=sysd["Oimag"]))
    driver.add_operator(operators.Operator(label="dipole", is_imaginary=False,
                                           ao_integrals=sysd["Oreal"]))
    driver.run(hamiltonian=hamiltonian, spin=spin_e, program=None, program_obj=None)
    return solver, driver, None


@pytest.mark.parametrize("name", ["water", "syn10", "lih"])
@pytest.mark.parametrize("tda_solver,ham", [(False, "rpa"), (False, "tda"), (True, "tda")])
@pytest.mark.parametrize("spin", ["singlet", "triplet"])
def test_excitation_energies_against_reference(golden_dir, name, tda_solver, ham, spin):
    from _td_cases import tag_of, td_system

    sysd = td_system(golden_dir, name)
    g = np.load(golden_dir / "td_golden.npz")
    tag = tag_of(name, tda_solver, ham, spin)
    solver, driver, prop = _run_td(sysd, tda_solver, ham, spin, ecd=True)
    ref_e = g[f"eigvals_{tag}"]
    assert solver.eigvals.shape == ref_e.shape and solver.eigvals.dtype == np.float64
    np.testing.assert_allclose(solver.eigvals, ref_e, rtol=0, atol=1e-10)
    # eigenvectors: unit columns, equal to the reference's up to sign
    ref_v = g[f"eigvecs_{tag}"]
    assert solver.eigvecs.shape == ref_v.shape
    np.testing.assert_allclose(np.linalg.norm(solver.eigvecs, axis=0), 1.0, rtol=0, atol=1e-12)
    overlap = np.abs(np.einsum("ds,ds->s", solver.eigvecs, ref_v))
    gaps = np.diff(ref_e)
    isolated = np.concatenate(([True], gaps > 1e-6)) & np.concatenate((gaps > 1e-6, [True]))
    np.testing.assert_allclose(overlap[isolated], 1.0, rtol=0, atol=1e-8)
    for op in solver.operators:
        np.testing.assert_allclose(np.abs(op.transition_moments)[isolated],
                                   np.abs(g[f"tmom_{op.label}_{tag}"])[isolated], rtol=0, atol=1e-8)
        np.testing.assert_allclose(op.oscillator_strengths[isolated],
                                   g[f"osc_{op.label}_{tag}"][isolated], rtol=0, atol=1e-8)
        np.testing.assert_allclose(op.total_oscillator_strengths[isolated],
                                   g[f"totosc_{op.label}_{tag}"][isolated], rtol=0, atol=1e-8)
    ref_rot = g[f"rotstr_{tag}"]
    np.testing.assert_allclose(prop.rotational_strengths_diplen[isolated], ref_rot[isolated], rtol=0,
                               atol=1e-7 * max(1.0, np.abs(ref_rot).max()))
    # the reports are text only; they must render
    assert "STATE  1" in prop.print_results_orca()
    assert "Root   1" in prop.print_results_nwchem()


def test_explicit_hessian_layout_matches_reference(golden_dir):
    """[[A, B], [-B, -A]] (RPA solver) and A (TDA solver) exactly as the reference lays them out."""
    from _td_cases import td_system

    sysd = td_system(golden_dir, "syn10")
    A, B = orc.rhf_ab(sysd["E"], sysd["tei"], sysd["ttype"], sysd["occ"][0], "rpa", "singlet")
    solver, _, _ = _run_td(sysd, False, "rpa", "singlet")
    want = np.block([[A, B], [-B, -A]])
    np.testing.assert_allclose(solver.explicit_hessian, want, rtol=0, atol=1e-12 * np.abs(want).max())
    solver, _, _ = _run_td(sysd, True, "tda", "singlet")
    np.testing.assert_allclose(solver.explicit_hessian, A, rtol=0, atol=1e-12 * np.abs(A).max())


def test_excitation_energies_larger_case_against_oracle():
    """n = 30, nocc = 6 (nov = 144, 288 x 288 reference matrix): all roots against scipy.linalg.eig
    on the oracle's explicit matrix; the transition table against the oracle's."""
    from pymolresponse_b200 import synthetic

    n, o = 30, 6
    case = synthetic.rhf_case(n, o)
    tei = orc.ao2mo_rhf_partial(case["I"], case["C"], o)
    sysd = {"C": case["C"], "E": case["E"], "occ": case["occupations"], "tei": tei,
            "ttype": "partial", "Oimag": synthetic.operator_integrals(n, 3, True),
            "Oreal": synthetic.operator_integrals(n, 3, False)}
    solver, driver, _ = _run_td(sysd, False, "rpa", "singlet")
    ev, vecs = orc.diagonalize_rhf(case["E"], tei, "partial", o, "rpa", "singlet")
    assert np.abs(ev.imag).max() == 0.0
    np.testing.assert_allclose(solver.eigvals, ev.real, rtol=0, atol=1e-10)
    grads = [orc.form_rhs(sysd["Oimag"], case["C"], case["occupations"], True)
             ["mo_integrals_ai_supervector_alph"][:, :, 0],
             orc.form_rhs(sysd["Oreal"], case["C"], case["occupations"], False)
             ["mo_integrals_ai_supervector_alph"][:, :, 0]]
    res = orc.transition_results(ev, vecs, grads, o, n - o)
    for op, r in zip(solver.operators, res):
        np.testing.assert_allclose(op.total_oscillator_strengths, r["total_oscillator_strengths"],
                                   rtol=0, atol=1e-8)
    assert solver.jacobi_sweeps <= 30


def test_unstable_reference_raises():
    """A + B not positive definite: the symmetric reduction does not apply; say so."""
    from pymolresponse_b200 import solvers, synthetic
    from pymolresponse_b200.core import AO2MOTransformationType, Hamiltonian, Spin

    n, o = 8, 2
    case = synthetic.rhf_case(n, o)
    tei = orc.ao2mo_rhf_partial(case["I"], case["C"], o)
    E = case["E"].copy()
    E[0][o, o] = E[0][0, 0] - 1.0     # a "virtual" below an occupied orbital: negative diagonal
    solver = solvers.ExactDiagonalizationSolver(case["C"], E, case["occupations"])
    solver.tei_mo = tei
    solver.tei_mo_type = AO2MOTransformationType.partial
    with pytest.raises(np.linalg.LinAlgError):
        solver.run(Hamiltonian.RPA, Spin.singlet, None, None)
