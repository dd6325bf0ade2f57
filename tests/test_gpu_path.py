"""GPU: the reference-facing API of pymolresponse_b200 against the golden vectors of the
unmodified reference and against the oracle on seeded inputs.

Tolerances (BASELINE.json north_star): index maps / Hessian layouts bit-exact; MO integrals and
Hessian elements <= 1e-12 * max|ref|; property tensors <= 1e-8 absolute (we hold 1e-10).
"""

import numpy as np
import pytest

from oracle import pmr_oracle as orc

pytestmark = pytest.mark.gpu

HAMS = ["rpa", "tda"]
SPINS = ["singlet", "triplet"]


def _enums():
    from pymolresponse_b200.core import AO2MOTransformationType, Hamiltonian, Spin

    return (AO2MOTransformationType,
            {"rpa": Hamiltonian.RPA, "tda": Hamiltonian.TDA},
            {"singlet": Spin.singlet, "triplet": Spin.triplet})


def _run(C, E, occ, tei, ttype, ops, freqs, ham, spin, solver_cls=None, **kw):
    from pymolresponse_b200 import cphf, operators, solvers

    T, H, S = _enums()
    solver = (solver_cls or solvers.ExactInv)(C, E, occ, **kw)
    solver.tei_mo = tei
    solver.tei_mo_type = T[ttype]
    driver = cphf.CPHF(solver)
    objs = []
    for o in ops:
        op = operators.Operator(label=o.get("label", "dipole"), is_imaginary=o["is_imaginary"],
                                is_spin_dependent=False, triplet=o.get("triplet", False),
                                ao_integrals=o["ao_integrals"])
        driver.add_operator(op)
        objs.append(op)
    driver.set_frequencies([float(f) for f in freqs])
    driver.run(hamiltonian=H[ham], spin=S[spin], program=None, program_obj=None)
    return solver, driver, objs


# ------------------------------------------------------------------------- Hessian blocks (K4/K5)
def test_hessian_blocks_bit_exact_vs_reference(golden_dir):
    from pymolresponse_b200 import explicit_equations_full as eqf
    from pymolresponse_b200 import explicit_equations_partial as eqp
    from pymolresponse_b200 import synthetic

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    E = synthetic.rhf_case(n, o, build_eri=False)["E"][0]
    ovov, oovv, full = g["ovov"], g["oovv"], g["full"]
    got = {
        "A_singlet_partial": eqp.form_rpa_a_matrix_mo_singlet_partial(E, ovov, oovv),
        "A_triplet_partial": eqp.form_rpa_a_matrix_mo_triplet_partial(E, oovv),
        "B_singlet_partial": eqp.form_rpa_b_matrix_mo_singlet_partial(ovov),
        "B_triplet_partial": eqp.form_rpa_b_matrix_mo_triplet_partial(ovov),
        "A_singlet_ss_partial": eqp.form_rpa_a_matrix_mo_singlet_ss_partial(E, ovov, oovv),
        "A_singlet_os_partial": eqp.form_rpa_a_matrix_mo_singlet_os_partial(ovov),
        "B_singlet_ss_partial": eqp.form_rpa_b_matrix_mo_singlet_ss_partial(ovov),
        "B_singlet_os_partial": eqp.form_rpa_b_matrix_mo_singlet_os_partial(ovov),
        "A_singlet_full": eqf.form_rpa_a_matrix_mo_singlet_full(E, full, o),
        "A_triplet_full": eqf.form_rpa_a_matrix_mo_triplet_full(E, full, o),
        "B_singlet_full": eqf.form_rpa_b_matrix_mo_singlet_full(full, o),
        "B_triplet_full": eqf.form_rpa_b_matrix_mo_triplet_full(full, o),
        "A_singlet_ss_full": eqf.form_rpa_a_matrix_mo_singlet_ss_full(E, full, o),
        "A_singlet_os_full": eqf.form_rpa_a_matrix_mo_singlet_os_full(full, o, o),
        "B_singlet_ss_full": eqf.form_rpa_b_matrix_mo_singlet_ss_full(full, o),
        "B_singlet_os_full": eqf.form_rpa_b_matrix_mo_singlet_os_full(full, o, o),
    }
    for key, val in got.items():
        assert isinstance(val, np.ndarray) and val.dtype == np.float64
        assert val.shape == g[key].shape, key
        assert np.array_equal(val, g[key]), key


def test_hessian_index_maps_with_integer_fill():
    """Layout contract: fill the tensors with exact integers that encode their own index and check
    every element of every block lands where the reference puts it (ragged sizes, tiles + edges)."""
    from pymolresponse_b200 import explicit_equations_full as eqf
    from pymolresponse_b200 import explicit_equations_partial as eqp

    o, v = 3, 37
    n = o + v
    ovov = np.arange(o * v * o * v, dtype=np.float64).reshape(o, v, o, v)
    oovv = 7.0 * np.arange(o * o * v * v, dtype=np.float64).reshape(o, o, v, v) + 1.0
    E = np.diag(np.arange(n, dtype=np.float64) * 3.0)
    assert np.array_equal(eqp.form_rpa_a_matrix_mo_singlet_partial(E, ovov, oovv),
                          orc.a_singlet_partial(E, ovov, oovv))
    assert np.array_equal(eqp.form_rpa_a_matrix_mo_triplet_partial(E, oovv),
                          orc.a_triplet_partial(E, oovv))
    assert np.array_equal(eqp.form_rpa_b_matrix_mo_singlet_partial(ovov), orc.b_singlet_partial(ovov))
    assert np.array_equal(eqp.form_rpa_b_matrix_mo_triplet_partial(ovov), orc.b_triplet_partial(ovov))
    assert np.array_equal(eqp.form_rpa_a_matrix_mo_singlet_ss_partial(E, ovov, oovv),
                          orc.a_singlet_ss_partial(E, ovov, oovv))
    assert np.array_equal(eqp.form_rpa_b_matrix_mo_singlet_ss_partial(ovov),
                          orc.b_singlet_ss_partial(ovov))
    # rectangular opposite-spin block (o_x, v_x) x (o_y, v_y)
    xxyy = np.arange(3 * 37 * 2 * 38, dtype=np.float64).reshape(3, 37, 2, 38)
    assert np.array_equal(eqp.form_rpa_a_matrix_mo_singlet_os_partial(xxyy), orc.a_singlet_os_partial(xxyy))
    assert np.array_equal(eqp.form_rpa_b_matrix_mo_singlet_os_partial(xxyy), orc.b_singlet_os_partial(xxyy))
    # full layout, n = 13, nocc = 4 / 3
    n = 13
    T = np.arange(n ** 4, dtype=np.float64).reshape(n, n, n, n)
    E = np.diag(np.arange(n, dtype=np.float64))
    assert np.array_equal(eqf.form_rpa_a_matrix_mo_singlet_full(E, T, 4), orc.a_singlet_full(E, T, 4))
    assert np.array_equal(eqf.form_rpa_a_matrix_mo_triplet_full(E, T, 4), orc.a_triplet_full(E, T, 4))
    assert np.array_equal(eqf.form_rpa_b_matrix_mo_singlet_full(T, 4), orc.b_singlet_full(T, 4))
    assert np.array_equal(eqf.form_rpa_b_matrix_mo_triplet_full(T, 4), orc.b_triplet_full(T, 4))
    assert np.array_equal(eqf.form_rpa_a_matrix_mo_singlet_ss_full(E, T, 4), orc.a_singlet_ss_full(E, T, 4))
    assert np.array_equal(eqf.form_rpa_b_matrix_mo_singlet_ss_full(T, 4), orc.b_singlet_ss_full(T, 4))
    assert np.array_equal(eqf.form_rpa_a_matrix_mo_singlet_os_full(T, 4, 3), orc.a_singlet_os_full(T, 4, 3))
    assert np.array_equal(eqf.form_rpa_b_matrix_mo_singlet_os_full(T, 4, 3), orc.b_singlet_os_full(T, 4, 3))


def test_hessian_blocks_empty_and_device_io():
    from pymolresponse_b200 import _native as nv
    from pymolresponse_b200 import explicit_equations_partial as eqp

    # no virtual orbitals: empty (0, 0) blocks, as NumPy would give
    E = np.diag(np.arange(3.0))
    A = eqp.form_rpa_a_matrix_mo_singlet_partial(E, np.zeros((3, 0, 3, 0)), np.zeros((3, 3, 0, 0)))
    assert A.shape == (0, 0)
    # CUDA tensors in -> CUDA tensor out
    r = np.random.default_rng(3)
    ovov = r.standard_normal((2, 5, 2, 5))
    B = eqp.form_rpa_b_matrix_mo_singlet_partial(nv.to_dev(ovov))
    assert nv.is_device_array(B)
    assert np.array_equal(nv.to_host(B), orc.b_singlet_partial(ovov))


# ------------------------------------------------------------------------- AO2MO (K1-K3)
def test_ao2mo_rhf_vs_reference(golden_dir):
    from pymolresponse_b200 import synthetic
    from pymolresponse_b200.ao2mo import AO2MO

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o)
    a = AO2MO(case["C"], case["occupations"], I=case["I"])
    a.perform_rhf_partial()
    assert len(a.tei_mo) == 2
    for got, key in zip(a.tei_mo, ("ovov", "oovv")):
        assert got.shape == g[key].shape and got.flags["C_CONTIGUOUS"]
        assert np.abs(got - g[key]).max() <= 1e-12 * np.abs(g[key]).max()
    a.perform_rhf_full()
    assert np.abs(a.tei_mo[0] - g["full"]).max() <= 1e-12 * np.abs(g["full"]).max()
    C = case["C"][0]
    t = AO2MO.transform(case["I"], C[:, :o], C[:, o:], C[:, :o], C[:, o:])
    assert np.abs(t - g["ovov"]).max() <= 1e-12 * np.abs(g["ovov"]).max()


def test_ao2mo_uhf_vs_reference(golden_dir):
    from pymolresponse_b200 import synthetic
    from pymolresponse_b200.ao2mo import AO2MO

    g = np.load(golden_dir / "synthetic_uhf_n9.npz")
    n, na, nb = int(g["n"]), int(g["na"]), int(g["nb"])
    case = synthetic.uhf_case(n, na, nb)
    a = AO2MO(case["C"], case["occupations"], I=case["I"])
    a.perform_uhf_partial()
    assert len(a.tei_mo) == 6
    for k, got in enumerate(a.tei_mo):
        ref = g[f"partial_{k}"]
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), k
    a.perform_uhf_full()
    assert len(a.tei_mo) == 4
    for k, got in enumerate(a.tei_mo):
        ref = g[f"full_{k}"]
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), k


@pytest.mark.parametrize("n,o,nu_block", [(24, 5, 0), (24, 5, 7), (33, 4, 16)])
def test_ao2mo_partial_vs_oracle_and_nu_slabs(n, o, nu_block):
    """Bigger/odd sizes against the oracle; nu-blocking and nu-sharded slabs give the same blocks."""
    from pymolresponse_b200 import _native as nv
    from pymolresponse_b200 import synthetic
    from pymolresponse_b200.ao2mo import AO2MO

    case = synthetic.rhf_case(n, o)
    # break the permutational symmetry: the transform must not assume any
    r = np.random.default_rng(1)
    I = case["I"] + 0.01 * r.standard_normal(case["I"].shape)
    ovov_ref, oovv_ref = orc.ao2mo_rhf_partial(I, case["C"], o)
    a = AO2MO(case["C"], case["occupations"], I=I, nu_block=nu_block)
    a.perform_rhf_partial()
    assert np.abs(a.tei_mo[0] - ovov_ref).max() <= 1e-12 * np.abs(ovov_ref).max()
    assert np.abs(a.tei_mo[1] - oovv_ref).max() <= 1e-12 * np.abs(oovv_ref).max()
    # two nu-slabs ("two ranks"): partial sums add up to the full blocks
    half = n // 2 + 1
    parts = []
    for lo, hi in ((0, half), (half, n)):
        s = AO2MO(case["C"], case["occupations"], I=np.ascontiguousarray(I[:, lo:hi]),
                  nu_range=(lo, hi), device_results=True)
        s.perform_rhf_partial()
        parts.append(s.tei_mo)
    ovov = nv.to_host(parts[0][0] + parts[1][0])
    oovv = nv.to_host(parts[0][1] + parts[1][1])
    assert np.abs(ovov - ovov_ref).max() <= 1e-12 * np.abs(ovov_ref).max()
    assert np.abs(oovv - oovv_ref).max() <= 1e-12 * np.abs(oovv_ref).max()


# ------------------------------------------------------------------------- RHS (K11)
def test_form_rhs_vs_reference(golden_dir):
    from pymolresponse_b200 import operators, synthetic
    from pymolresponse_b200.indices import form_indices_from_occupations

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o, build_eri=False)
    idx = form_indices_from_occupations(case["occupations"])
    op = operators.Operator(label="dipole", is_imaginary=False,
                            ao_integrals=synthetic.operator_integrals(n, 3, False))
    op.form_rhs(case["C"], case["occupations"], idx)
    assert op.mo_integrals_ai_alph.shape == g["rhs0_ai"].shape
    assert op.mo_integrals_ai_supervector_alph.shape == g["rhs0_super"].shape
    assert np.abs(op.mo_integrals_ai_alph - g["rhs0_ai"]).max() <= 1e-14
    assert np.abs(op.mo_integrals_ai_supervector_alph - g["rhs0_super"]).max() <= 1e-14
    op1 = operators.Operator(label="angmom", is_imaginary=True,
                             ao_integrals=synthetic.operator_integrals(n, 3, True))
    op1.form_rhs(case["C"], case["occupations"], idx)
    assert np.abs(op1.mo_integrals_ai_supervector_alph - g["rhs1_super"]).max() <= 1e-14
    # triplet masking + spin-orbit prefactor against the oracle (operators.py:39-40, 91-98)
    occ = (4, 6, 3, 7)
    Cu = np.stack((case["C"][0], case["C"][0][:, ::-1].copy()))
    O = synthetic.operator_integrals(n, 2, True)
    op2 = operators.Operator(label="spinorb1", is_imaginary=True, triplet=True, ao_integrals=O)
    op2.form_rhs(Cu, occ, form_indices_from_occupations(occ))
    ref = orc.form_rhs(O, Cu, occ, is_imaginary=True, triplet=True, hsofac=op2.hsofac)
    for tag in ("alph", "beta"):
        key = f"mo_integrals_ai_supervector_{tag}"
        assert np.abs(getattr(op2, key) - ref[key]).max() <= 1e-16


# ------------------------------------------------------------------------- full path: water (config 1)
WATER_LITERALS = {
    ("rpa", "singlet"): [[7.93556221, 3.06821077, 0.05038621], [7.94312371, 3.07051688, 0.05054685],
                         [8.00414009, 3.08913608, 0.05186977], [8.1290378, 3.12731363, 0.05473482]],
    ("rpa", "triplet"): [[26.59744305, 18.11879557, 0.07798969], [26.68282287, 18.19390051, 0.07837521],
                         [27.38617401, 18.81922578, 0.08160226], [28.91067234, 20.21670386, 0.08892512]],
    ("tda", "singlet"): [[8.89855952, 4.00026556, 0.0552774], [8.90690928, 4.00298342, 0.05545196],
                         [8.97427725, 4.02491517, 0.05688918], [9.11212633, 4.06981937, 0.05999934]],
    ("tda", "triplet"): [[14.64430714, 8.80921432, 0.06859496], [14.68168443, 8.83562647, 0.0689291],
                         [14.98774296, 9.0532224, 0.07172414], [15.63997724, 9.52504267, 0.07805428]],
}


@pytest.mark.parametrize("ham", HAMS)
@pytest.mark.parametrize("spin", SPINS)
def test_water_sto3g_polarizability(golden_dir, ham, spin):
    """Config 1: the reference's tests/test_final_result.py re-hosted on the GPU path."""
    from pymolresponse_b200 import solvers

    g = np.load(golden_dir / "water_sto3g.npz")
    ops = [{"ao_integrals": g["dipole"], "is_imaginary": False}]
    for cls, tag in ((solvers.ExactInv, "inv"), (solvers.ExactInvCholesky, "chol")):
        solver, driver, objs = _run(g["C"], g["E"], (5, 2, 5, 2), (g["TEI_MO"],), "full", ops,
                                    g["frequencies"], ham, spin, solver_cls=cls)
        assert len(driver.results) == 4
        for f in range(4):
            assert driver.results[f].shape == (3, 3)
            np.testing.assert_allclose(driver.results[f], np.diag(WATER_LITERALS[(ham, spin)][f]),
                                       rtol=0, atol=1e-8)
        key = f"{ham}_{spin}_{tag}"
        np.testing.assert_allclose(np.stack(driver.results), g[f"results_{key}"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(np.stack(driver.uncoupled_results), g[f"uncoupled_{key}"], rtol=0,
                                   atol=1e-12)
        assert objs[0].rspvecs_alph[0].shape == (3, 20, 1)
        np.testing.assert_allclose(np.stack(objs[0].rspvecs_alph), g[f"rspvecs_{key}"], rtol=0,
                                   atol=1e-10)
        assert len(objs[0].rspvecs_beta) == 0
    # lazily materialised explicit Hessian of the last frequency: bit-exact layout
    solver, _, _ = _run(g["C"], g["E"], (5, 2, 5, 2), (g["TEI_MO"],), "full", ops, g["frequencies"],
                        ham, spin)
    assert np.array_equal(solver.explicit_hessian, g[f"hessian_last_{ham}_{spin}_inv"])


# ------------------------------------------------------------------------- full path: synthetic RHF
@pytest.mark.parametrize("ham", HAMS)
@pytest.mark.parametrize("spin", SPINS)
@pytest.mark.parametrize("ttype", ["partial", "full"])
def test_synthetic_rhf_mixed_operators(golden_dir, ham, spin, ttype):
    from pymolresponse_b200 import synthetic

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o, build_eri=False)
    ops = [{"label": "dipole", "ao_integrals": synthetic.operator_integrals(n, 3, False),
            "is_imaginary": False},
           {"label": "angmom", "ao_integrals": synthetic.operator_integrals(n, 3, True),
            "is_imaginary": True}]
    tei = (g["ovov"], g["oovv"]) if ttype == "partial" else (g["full"],)
    solver, driver, objs = _run(case["C"], case["E"], case["occupations"], tei, ttype, ops,
                                g["frequencies"], ham, spin)
    tag = f"{ham}_{spin}_{ttype}"
    np.testing.assert_allclose(np.stack(driver.results), g[f"results_{tag}"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(np.stack(driver.uncoupled_results), g[f"uncoupled_{tag}"], rtol=0,
                               atol=1e-12)
    for k in (0, 1):
        ref = g[f"rspvecs{k}_{tag}"]
        got = np.stack(objs[k].rspvecs_alph)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    if ttype == "partial":
        assert np.array_equal(solver.explicit_hessian, g[f"hessian_last_{tag}"])
        ref_inv = g[f"hessian_inv_last_{tag}"]
        assert np.abs(solver.explicit_hessian_inv - ref_inv).max() <= 1e-10 * np.abs(ref_inv).max()


def test_rhf_cholesky_solver_and_api_quirks(golden_dir):
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200.core import Hamiltonian, Spin

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o, build_eri=False)
    ops = [{"ao_integrals": synthetic.operator_integrals(n, 3, False), "is_imaginary": False},
           {"ao_integrals": synthetic.operator_integrals(n, 3, True), "is_imaginary": True}]
    _, driver, _ = _run(case["C"], case["E"], case["occupations"], (g["ovov"], g["oovv"]), "partial",
                        ops, g["frequencies"], "rpa", "singlet", solver_cls=solvers.ExactInvCholesky)
    np.testing.assert_allclose(np.stack(driver.results), g["results_rpa_singlet_partial_chol"],
                               rtol=0, atol=1e-10)
    # quirks the drop-in keeps: frequencies must be float (solvers.py:173); no program and no
    # tei_mo is a RuntimeError (solvers.py:473-475); results default to the static case
    s = solvers.ExactInv(case["C"], case["E"], case["occupations"])
    d = cphf.CPHF(s)
    d.add_operator(operators.Operator(label="dipole", is_imaginary=False,
                                      ao_integrals=ops[0]["ao_integrals"]))
    with pytest.raises(RuntimeError):
        d.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    s.tei_mo = (g["ovov"], g["oovv"])
    d.set_frequencies([0])
    with pytest.raises(AssertionError):
        d.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    s2 = solvers.ExactInv(case["C"], case["E"], case["occupations"])
    s2.tei_mo = (g["ovov"], g["oovv"])
    d2 = cphf.CPHF(s2)
    d2.add_operator(operators.Operator(label="dipole", is_imaginary=False,
                                       ao_integrals=ops[0]["ao_integrals"]))
    assert not hasattr(d2, "results")
    d2.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    assert d2.frequencies == [0.0]
    np.testing.assert_allclose(d2.results[0], g["results_rpa_singlet_partial"][0][:3, :3], rtol=0,
                               atol=1e-10)


def test_rhf_above_first_pole_lu_fallback(golden_dir):
    """w above the first pole: G(w) is indefinite.  ExactInv must still return inv(G) rhs (pivoted
    LU on the device); ExactInvCholesky must raise LinAlgError like scipy.linalg.cholesky."""
    from pymolresponse_b200 import solvers, synthetic

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o, build_eri=False)
    Or = synthetic.operator_integrals(n, 3, False)
    ops = [{"ao_integrals": Or, "is_imaginary": False}]
    freqs = [0.1, 1.9]
    tei = (g["ovov"], g["oovv"])
    ref = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial", ops, freqs)
    solver, driver, _ = _run(case["C"], case["E"], case["occupations"], tei, "partial", ops, freqs,
                             "rpa", "singlet")
    assert solver.solve_stats["lu_fallback"] >= 1
    for f in range(2):
        np.testing.assert_allclose(driver.results[f], ref["results"][f], rtol=0, atol=1e-9)
    with pytest.raises(np.linalg.LinAlgError):
        _run(case["C"], case["E"], case["occupations"], tei, "partial", ops, freqs, "rpa", "singlet",
             solver_cls=solvers.ExactInvCholesky)


def test_rhf_custom_inv_func(golden_dir):
    from pymolresponse_b200 import synthetic

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o, build_eri=False)
    ops = [{"ao_integrals": synthetic.operator_integrals(n, 3, False), "is_imaginary": False},
           {"ao_integrals": synthetic.operator_integrals(n, 3, True), "is_imaginary": True}]
    calls = []

    def my_inv(G):
        calls.append(G.shape)
        return np.linalg.pinv(G)

    _, driver, _ = _run(case["C"], case["E"], case["occupations"], (g["ovov"], g["oovv"]), "partial",
                        ops, g["frequencies"], "rpa", "singlet", inv_func=my_inv)
    assert len(calls) == len(g["frequencies"])
    np.testing.assert_allclose(np.stack(driver.results), g["results_rpa_singlet_partial"], rtol=0,
                               atol=1e-9)


# ------------------------------------------------------------------------- full path: synthetic UHF
@pytest.mark.parametrize("ham", HAMS)
@pytest.mark.parametrize("spin", SPINS)
@pytest.mark.parametrize("ttype", ["partial", "full"])
def test_synthetic_uhf_mixed_operators(golden_dir, ham, spin, ttype):
    from pymolresponse_b200 import synthetic

    g = np.load(golden_dir / "synthetic_uhf_n9.npz")
    n, na, nb = int(g["n"]), int(g["na"]), int(g["nb"])
    case = synthetic.uhf_case(n, na, nb, build_eri=False)
    ops = [{"label": "dipole", "ao_integrals": synthetic.operator_integrals(n, 3, False),
            "is_imaginary": False},
           {"label": "angmom", "ao_integrals": synthetic.operator_integrals(n, 3, True),
            "is_imaginary": True}]
    tei = tuple(g[f"{ttype}_{k}"] for k in range(6 if ttype == "partial" else 4))
    solver, driver, objs = _run(case["C"], case["E"], case["occupations"], tei, ttype, ops,
                                g["frequencies"], ham, spin)
    tag = f"{ham}_{spin}_{ttype}"
    np.testing.assert_allclose(np.stack(driver.results), g[f"results_{tag}"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(np.stack(driver.uncoupled_results), g[f"uncoupled_{tag}"], rtol=0,
                               atol=1e-12)
    for k in (0, 1):
        for sp in ("alph", "beta"):
            ref = g[f"rspvecs{k}_{sp}_{tag}"]
            got = np.stack(getattr(objs[k], f"rspvecs_{sp}"))
            assert got.shape == ref.shape
            assert np.abs(got - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    if tag == "rpa_singlet_partial":
        G4 = solver.explicit_hessian
        assert isinstance(G4, tuple) and len(G4) == 4
        for q in range(4):
            assert np.array_equal(G4[q], g[f"hessian_last_{q}"])
        solver.invert_explicit_hessian()
        assert np.abs(solver.left_alph - g["left_alph_last"]).max() <= 1e-10 * np.abs(g["left_alph_last"]).max()
        assert np.abs(solver.left_beta - g["left_beta_last"]).max() <= 1e-10 * np.abs(g["left_beta_last"]).max()


def test_uhf_cholesky_solver(golden_dir):
    from pymolresponse_b200 import solvers, synthetic

    g = np.load(golden_dir / "synthetic_uhf_n9.npz")
    n, na, nb = int(g["n"]), int(g["na"]), int(g["nb"])
    case = synthetic.uhf_case(n, na, nb, build_eri=False)
    ops = [{"ao_integrals": synthetic.operator_integrals(n, 3, False), "is_imaginary": False},
           {"ao_integrals": synthetic.operator_integrals(n, 3, True), "is_imaginary": True}]
    tei = tuple(g[f"partial_{k}"] for k in range(6))
    _, driver, _ = _run(case["C"], case["E"], case["occupations"], tei, "partial", ops,
                        g["frequencies"], "rpa", "singlet", solver_cls=solvers.ExactInvCholesky)
    np.testing.assert_allclose(np.stack(driver.results), g["results_rpa_singlet_partial_chol"],
                               rtol=0, atol=1e-10)


# ------------------------------------------------------------------------- LiH disk fixtures
@pytest.mark.parametrize("case,uhf,thresh", [("r_lih_hf_sto-3g", False, 5.0e-3),
                                             ("u_lih_cation_hf_sto-3g", True, 1.0e-1)])
def test_lih_dalton_fixtures(golden_dir, case, uhf, thresh):
    """tests/test_runners.py re-hosted: DALTON values with the reference's thresholds, the
    reference's own numbers to 1e-10, operator-interchange symmetry (test_calculators.py:87-91)."""
    from pymolresponse_b200 import cphf, solvers
    from pymolresponse_b200.core import AO2MOTransformationType, Hamiltonian, Spin
    from pymolresponse_b200.interfaces.dalton.utils import dalton_label_to_operator

    g = np.load(golden_dir / f"{case}.npz")
    occ = tuple(int(x) for x in g["occupations"])
    if uhf:
        E = np.stack((np.diag(g["moene"][:, 0]), np.diag(g["moene"][:, 1])), axis=0)
        tei = tuple(g[k] for k in ("moints_iajb_aaaa", "moints_iajb_aabb", "moints_iajb_bbaa",
                                   "moints_iajb_bbbb", "moints_ijab_aaaa", "moints_ijab_bbbb"))
    else:
        E = np.diag(g["moene"][:, 0])[None]
        tei = (g["moints_iajb_aaaa"], g["moints_ijab_aaaa"])
    idx = [k for k in range(len(g["ref_value"])) if not np.isnan(g["reference_value"][k])]
    checked = 0
    for k in idx[3::41]:
        solver = solvers.ExactInv(g["C"], E, occ)
        solver.tei_mo = tei
        solver.tei_mo_type = AO2MOTransformationType.partial
        driver = cphf.CPHF(solver)
        for lab in (str(g["ref_label_1"][k]), str(g["ref_label_2"][k])):
            op = dalton_label_to_operator(lab)
            op.ao_integrals = g[f"operator_mn_{op.label}"][op.slice_idx][None]
            driver.add_operator(op)
        driver.set_frequencies([float(g["ref_frequency"][k])])
        driver.run(hamiltonian=Hamiltonian[str(g["ref_hamiltonian"][k]).upper()],
                   spin=Spin[str(g["ref_spin"][k])], program=None, program_obj=None)
        res = driver.results[0]
        assert res.shape == (2, 2)
        assert abs(abs(res[1, 0]) - abs(res[0, 1])) < 1e-12
        assert abs(res[1, 0] - g["reference_value"][k]) < 1e-10 * max(1.0, abs(g["reference_value"][k]))
        assert abs(g["ref_value"][k]) - abs(res[1, 0]) < thresh
        checked += 1
    assert checked >= 90


# ------------------------------------------------------------------------- iterative path (K10)
def test_iterative_solver_vs_reference(golden_dir):
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200.core import Hamiltonian, Spin
    from pymolresponse_b200.integrals import JKDense

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o)
    it = solvers.IterativeLinEqSolver(case["C"], case["E"], case["occupations"], JKDense(case["I"]),
                                      maxiter=200, conv=1e-24)
    drv = cphf.CPHF(it)
    op = operators.Operator(label="dipole", is_imaginary=False,
                            ao_integrals=synthetic.operator_integrals(n, 3, False))
    drv.add_operator(op)
    drv.set_frequencies([0.0, 0.05])
    drv.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    assert all(k > 0 for k in it.iterations)
    np.testing.assert_allclose(np.stack(op.rspvecs_alph), g["iterative_rspvecs"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(np.stack(drv.results), g["iterative_results"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(drv.results[0], g["results_rpa_singlet_partial"][0][:3, :3], rtol=0,
                               atol=1e-8)


# ------------------------------------------------------------------------- property wrappers
def test_polarizability_and_magnetizability_wrappers():
    """End to end from AO integrals (AO2MO on the GPU inside solver.run) through the wrappers,
    against the oracle; n = 40, nocc = 6 (nov = 204: several Cholesky panels)."""
    from pymolresponse_b200 import cphf, integrals, solvers, synthetic
    from pymolresponse_b200.core import Hamiltonian, Program, Spin
    from pymolresponse_b200.properties.electric import Polarizability
    from pymolresponse_b200.properties.magnetic import Magnetizability

    n, o = 40, 6
    case = synthetic.rhf_case(n, o)
    Or, Oi = synthetic.operator_integrals(n, 3, False), synthetic.operator_integrals(n, 3, True)
    gen = integrals.IntegralsArrays({"dipole": Or, "angmom_common_gauge": Oi})
    freqs = [0.0, 0.1, 0.2, 0.3]
    tei = orc.ao2mo_rhf_partial(case["I"], case["C"], o)
    ref_p = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial",
                         [{"ao_integrals": Or, "is_imaginary": False}], freqs)
    ref_m = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial",
                         [{"ao_integrals": Oi, "is_imaginary": True}], [0.0])

    class Prog:  # stands for "a program object that can hand out AO ERIs"
        ao_eri = case["I"]

    solver = solvers.ExactInv(case["C"], case["E"], case["occupations"])
    pol = Polarizability(Program.Arrays, gen, cphf.CPHF(solver), frequencies=freqs)
    pol.form_operators()
    solver.form_tei_mo(Program.Arrays, Prog(), solver.tei_mo_type)
    assert np.abs(solver.tei_mo[0] - tei[0]).max() <= 1e-12 * np.abs(tei[0]).max()
    pol.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet)
    pol.form_results()
    assert len(pol.polarizabilities) == 4
    for f in range(4):
        np.testing.assert_allclose(pol.polarizabilities[f], ref_p["results"][f], rtol=0, atol=1e-9)
        np.testing.assert_allclose(pol.polarizabilities[f], pol.polarizabilities[f].T, rtol=0,
                                   atol=1e-11)
    solver_m = solvers.ExactInvCholesky(case["C"], case["E"], case["occupations"])
    solver_m.tei_mo = solver.tei_mo
    mag = Magnetizability(Program.Arrays, gen, cphf.CPHF(solver_m))
    mag.form_operators()
    mag.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet)
    mag.form_results()
    np.testing.assert_allclose(mag.magnetizability, ref_m["results"][0] / 4, rtol=0, atol=1e-9)


def test_ao2mo_from_factor_matches_dense_route():
    """MO blocks straight from the factor B[P, mu nu] (no n^4 tensor) against the
    oracle's dense transforms of I = B^T B; RHF/UHF, partial/full, and through Solver.form_tei_mo."""
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200.ao2mo import AO2MO
    from pymolresponse_b200.core import Hamiltonian, Program, Spin

    def close(got, want):
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())

    n, o = 21, 4
    case = synthetic.rhf_case(n, o)
    a = AO2MO(case["C"], case["occupations"], df_factor=case["B"])
    a.perform_rhf_partial()
    for got, want in zip(a.tei_mo, orc.ao2mo_rhf_partial(case["I"], case["C"], o)):
        close(got, want)
    a.perform_rhf_full()
    close(a.tei_mo[0], orc.ao2mo_rhf_full(case["I"], case["C"])[0])

    na, nb = 5, 3
    ucase = synthetic.uhf_case(n, na, nb)
    u = AO2MO(ucase["C"], ucase["occupations"], df_factor=ucase["B"])
    u.perform_uhf_partial()
    want = orc.ao2mo_uhf_partial(ucase["I"], ucase["C"], ucase["occupations"])
    assert len(u.tei_mo) == 6
    for g_, w_ in zip(u.tei_mo, want):
        close(g_, w_)
    u.perform_uhf_full()
    for g_, w_ in zip(u.tei_mo, orc.ao2mo_uhf_full(ucase["I"], ucase["C"])):
        close(g_, w_)

    class Prog:
        df_factor = case["B"]

    Or = synthetic.operator_integrals(n, 3, False)
    ref = orc.run_cphf(case["C"], case["E"], case["occupations"],
                       orc.ao2mo_rhf_partial(case["I"], case["C"], o), "partial",
                       [{"ao_integrals": Or, "is_imaginary": False}], [0.0, 0.1])
    solver = solvers.ExactInv(case["C"], case["E"], case["occupations"])
    driver = cphf.CPHF(solver)
    driver.add_operator(operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or))
    driver.set_frequencies([0.0, 0.1])
    driver.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=Program.Arrays,
               program_obj=Prog())
    for f in range(2):
        np.testing.assert_allclose(driver.results[f], ref["results"][f], rtol=0, atol=1e-9)


def test_ord_and_gtensor_wrappers():
    """Mixed imaginary/real multi-operator batches through the ORD and g-tensor wrappers (reference
    properties/optrot.py:13-96, properties/magnetic.py:57-190) against the oracle's result blocks."""
    from pymolresponse_b200 import cphf, integrals, solvers, synthetic
    from pymolresponse_b200.core import Hamiltonian, Program, Spin
    from pymolresponse_b200.properties.magnetic import ElectronicGTensor
    from pymolresponse_b200.properties.optrot import ORD

    n, o = 24, 5
    case = synthetic.rhf_case(n, o)
    Oi = synthetic.operator_integrals(n, 3, True)
    Or = synthetic.operator_integrals(n, 3, False)
    Ov = synthetic.operator_integrals(n, 3, True, seed=77)
    gen = integrals.IntegralsArrays({"dipole": Or, "angmom_common_gauge": Oi, "dipvel": Ov})
    tei = orc.ao2mo_rhf_partial(case["I"], case["C"], o)
    freqs = [0.0, 0.12]
    for do_dipvel in (False, True):
        ops = [{"ao_integrals": Oi, "is_imaginary": True}, {"ao_integrals": Or, "is_imaginary": False}]
        if do_dipvel:
            ops.append({"ao_integrals": Ov, "is_imaginary": True})
        ref = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial", ops, freqs)
        solver = solvers.ExactInv(case["C"], case["E"], case["occupations"])
        solver.tei_mo = tei
        ord_ = ORD(Program.Arrays, gen, cphf.CPHF(solver), frequencies=freqs, do_dipvel=do_dipvel)
        ord_.form_operators()
        ord_.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet)
        ord_.form_results()
        for f in range(len(freqs)):
            np.testing.assert_allclose(ord_.driver.results[f], ref["results"][f], rtol=0, atol=1e-9)
            np.testing.assert_allclose(ord_.polarizabilities[f], ref["results"][f][3:6, 3:6],
                                       rtol=0, atol=1e-9)
            np.testing.assert_allclose(ord_.polarizabilities_lenmag[f], ref["results"][f][0:3, 3:6],
                                       rtol=0, atol=1e-9)

    # g-tensor: UHF doublet, three imaginary operators in one static batch
    na, nb = 5, 4
    ucase = synthetic.uhf_case(n, na, nb)
    Oso = synthetic.operator_integrals(n, 3, True, seed=78)
    gen_g = integrals.IntegralsArrays({"angmom_common_gauge": Oi, "spinorb": Oso,
                                       "spinorb_eff": 0 * Oso})
    utei = orc.ao2mo_uhf_partial(ucase["I"], ucase["C"], ucase["occupations"])
    # operators whose label contains "spinorb" carry the factor alpha^2 / 4 (operators.py:44-45)
    from pymolresponse_b200.constants import alpha

    hso = alpha**2 / 4
    ops = [{"ao_integrals": Oi, "is_imaginary": True},
           {"ao_integrals": Oso, "is_imaginary": True, "hsofac": hso},
           {"ao_integrals": 0 * Oso, "is_imaginary": True, "hsofac": hso}]
    ref = orc.run_cphf(ucase["C"], ucase["E"], ucase["occupations"], utei, "partial", ops, [0.0])
    usolver = solvers.ExactInv(ucase["C"], ucase["E"], ucase["occupations"])
    usolver.tei_mo = utei
    gt = ElectronicGTensor(Program.Arrays, gen_g, cphf.CPHF(usolver), gauge_origin=[0.0, 0.0, 0.0],
                           nelec=(na, nb))
    gt.form_operators()
    gt.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet)
    gt.form_results()
    want = ref["results"][0][0:3, 3:6] / (0.5 * (na - nb))
    np.testing.assert_allclose(gt.g_oz_soc_1, want, rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.sort(gt.g_oz_soc_1_eig.real),
                               np.sort(np.sqrt(np.linalg.eigvals(want.T @ want)).real), atol=1e-12)
    assert np.abs(want).max() > 1e-9


def test_uhf_end_to_end_from_ao_integrals():
    """Config-4 shaped small case: UHF (5 alpha, 4 beta), AO2MO + joint Cholesky on the GPU."""
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200.core import AO2MOTransformationType, Hamiltonian, Program, Spin

    n, na, nb = 26, 5, 4
    case = synthetic.uhf_case(n, na, nb)
    Or = synthetic.operator_integrals(n, 3, False)
    tei = orc.ao2mo_uhf_partial(case["I"], case["C"], case["occupations"])
    ref = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial",
                       [{"ao_integrals": Or, "is_imaginary": False}], [0.0, 0.15])
    solver = solvers.ExactInv(case["C"], case["E"], case["occupations"])
    driver = cphf.CPHF(solver)
    op = operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or)
    driver.add_operator(op)
    driver.set_frequencies([0.0, 0.15])
    driver.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=Program.Arrays,
               program_obj=case["I"])
    assert solver.tei_mo_type == AO2MOTransformationType.partial and len(solver.tei_mo) == 6
    for f in range(2):
        np.testing.assert_allclose(driver.results[f], ref["results"][f], rtol=0, atol=1e-9)
        assert np.abs(op.rspvecs_alph[f] - ref["rspvecs_alph"][0][f]).max() <= 1e-9
        assert np.abs(op.rspvecs_beta[f] - ref["rspvecs_beta"][0][f]).max() <= 1e-9


# ------------------------------------------------------------------------- larger sizes: properties
def test_larger_rhf_size_independent_properties():
    """n = 96, nocc = 12 (nov = 1008): too slow for repeated oracle runs inside the GPU suite, so
    check what the domain guarantees: (ia|jb) pair symmetry, M/K symmetry, the residual of the
    solved equations G(w)[X;Y] = rhs formed with the bit-exact A and B blocks, the symmetry of the
    polarizability tensor, and agreement of the Cholesky path with the pivoted-LU path."""
    from pymolresponse_b200 import _engine as eng
    from pymolresponse_b200 import _native as nv
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200 import explicit_equations_partial as eqp
    from pymolresponse_b200.ao2mo import AO2MO
    from pymolresponse_b200.core import Hamiltonian, Spin

    n, o = 96, 12
    v = n - o
    case = synthetic.rhf_case(n, o)
    a = AO2MO(case["C"], case["occupations"], I=case["I"])
    a.perform_rhf_partial()
    ovov, oovv = a.tei_mo
    nov = o * v
    assert np.abs(ovov.reshape(nov, nov) - ovov.reshape(nov, nov).T).max() <= 1e-13
    assert np.abs(oovv - oovv.transpose(1, 0, 2, 3)).max() <= 1e-13
    Or = synthetic.operator_integrals(n, 3, False)
    freqs = [0.0, 0.2]
    solver = solvers.ExactInv(case["C"], case["E"], case["occupations"])
    solver.tei_mo = (ovov, oovv)
    driver = cphf.CPHF(solver)
    op = operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or)
    driver.add_operator(op)
    driver.set_frequencies(freqs)
    driver.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    A = eqp.form_rpa_a_matrix_mo_singlet_partial(case["E"][0], ovov, oovv)
    B = eqp.form_rpa_b_matrix_mo_singlet_partial(ovov)
    rhs = op.mo_integrals_ai_supervector_alph[:, :, 0]
    for f, w in enumerate(freqs):
        G = np.block([[A, B], [B, A]]) - w * np.diag(np.concatenate((np.ones(nov), -np.ones(nov))))
        X = op.rspvecs_alph[f][:, :, 0]
        resid = X @ G.T - rhs
        assert np.abs(resid).max() <= 1e-11 * max(1.0, np.abs(rhs).max())
        assert np.abs(driver.results[f] - driver.results[f].T).max() <= 1e-10
    # pivoted LU on the same G agrees with the batched-Cholesky path
    sysm = solver._prepare(Hamiltonian.RPA, Spin.singlet, want_ab=True)
    g_rows = nv.to_dev(rhs[:, :nov])
    xy_lu = nv.to_host(sysm.solve_lu(g_rows, [-1, -1, -1], 0.2))
    assert np.abs(xy_lu - op.rspvecs_alph[1][:, :, 0]).max() <= 1e-10
    del eng


# ------------------------------------------------------------------------- BASELINE configs 2 and 4
def test_config2_shaped_cholesky_sweep():
    """BASELINE config 2 needs pyscf for H2O/aug-cc-pVTZ integrals (absent); same shape with the
    synthetic generator: nbasis=92, nocc=5, 8 frequencies, ExactInvCholesky, from AO integrals."""
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200.core import Hamiltonian, Program, Spin

    n, o = 92, 5
    case = synthetic.rhf_case(n, o)
    Or = synthetic.operator_integrals(n, 3, False)
    freqs = [float(w) for w in np.linspace(0.0, 0.35, 8)]
    tei = orc.ao2mo_rhf_partial(case["I"], case["C"], o)
    ref = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial",
                       [{"ao_integrals": Or, "is_imaginary": False}], freqs, cholesky=True)
    solver = solvers.ExactInvCholesky(case["C"], case["E"], case["occupations"])
    driver = cphf.CPHF(solver)
    op = operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or)
    driver.add_operator(op)
    driver.set_frequencies(freqs)
    driver.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=Program.Arrays,
               program_obj=case["I"])
    for k in range(2):
        assert np.abs(solver.tei_mo[k] - tei[k]).max() <= 1e-12 * np.abs(tei[k]).max()
    for f in range(8):
        np.testing.assert_allclose(driver.results[f], ref["results"][f], rtol=0, atol=1e-9)
        assert np.abs(op.rspvecs_alph[f] - ref["rspvecs_alph"][0][f]).max() <= 1e-9
        np.testing.assert_allclose(driver.uncoupled_results[f], ref["uncoupled_results"][f], rtol=0,
                                   atol=1e-11)


def test_config4_shaped_uhf_static():
    """BASELINE config 4 shape at 1/8 scale: UHF nbasis=48 (6 alpha, 5 beta), odd beta dimensions
    (8-byte-aligned copy paths, generic assembly kernel), static polarizability, joint Cholesky."""
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200.core import Hamiltonian, Program, Spin

    n, na, nb = 48, 6, 5
    case = synthetic.uhf_case(n, na, nb)
    Or = synthetic.operator_integrals(n, 3, False)
    tei = orc.ao2mo_uhf_partial(case["I"], case["C"], case["occupations"])
    ref = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial",
                       [{"ao_integrals": Or, "is_imaginary": False}], [0.0])
    for cls in (solvers.ExactInv, solvers.ExactInvCholesky):
        solver = cls(case["C"], case["E"], case["occupations"])
        driver = cphf.CPHF(solver)
        op = operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or)
        driver.add_operator(op)
        driver.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=Program.Arrays,
                   program_obj=case["I"])
        for k in range(6):
            assert np.abs(solver.tei_mo[k] - tei[k]).max() <= 1e-12 * np.abs(tei[k]).max(), k
        np.testing.assert_allclose(driver.results[0], ref["results"][0], rtol=0, atol=1e-9)
        assert np.abs(op.rspvecs_alph[0] - ref["rspvecs_alph"][0][0]).max() <= 1e-9
        assert np.abs(op.rspvecs_beta[0] - ref["rspvecs_beta"][0][0]).max() <= 1e-9


def test_config5_full_size_properties():
    """BASELINE config 5 at FULL size (nbasis=256, nocc=32: 34.4 GB AO tensor built on the device,
    nov = 7168), checked through size-independent properties: pair symmetry of the MO integrals,
    residual of G(w)[X;Y] = rhs with the bit-exact A/B blocks (torch is only the checker here),
    symmetry of the polarizability, linearity in the operator, real/imaginary operator in one batch."""
    import torch

    from pymolresponse_b200 import _native as nv
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200 import explicit_equations_partial as eqp
    from pymolresponse_b200.ao2mo import AO2MO
    from pymolresponse_b200.core import Hamiltonian, Spin

    free, _ = torch.cuda.mem_get_info()
    if free < 80 * 2**30:
        pytest.skip("needs ~80 GB of free HBM")
    n, o = 256, 32
    v, nov = n - o, o * (n - o)
    case = synthetic.rhf_case(n, o, build_eri=False)
    Bd = torch.from_numpy(case["B"]).cuda()
    Bf = Bd.reshape(n, n * n)
    I = torch.empty(n, n, n, n, dtype=torch.float64, device="cuda")
    for mu in range(n):
        I[mu] = (Bd[:, mu, :].t() @ Bf).reshape(n, n, n)
    del Bd, Bf
    a = AO2MO(case["C"], case["occupations"], I=I, device_results=True)
    a.perform_rhf_partial()
    ovov, oovv = a.tei_mo
    del I, a
    torch.cuda.empty_cache()
    m = ovov.reshape(nov, nov)
    assert float((m - m.t()).abs().max()) <= 1e-13
    assert float((oovv - oovv.transpose(0, 1)).abs().max()) <= 1e-13
    del m
    Or = synthetic.operator_integrals(n, 3, False)
    Oi = synthetic.operator_integrals(n, 3, True)
    freqs = [0.0, 0.3, 0.63]
    solver = solvers.ExactInv(case["C"], case["E"], case["occupations"])
    solver.device_results = True
    solver.tei_mo = (ovov, oovv)
    driver = cphf.CPHF(solver)
    ops = [operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or),
           operators.Operator(label="angmom", is_imaginary=True, ao_integrals=Oi),
           operators.Operator(label="dipole2", is_imaginary=False, ao_integrals=2.5 * Or)]
    for op in ops:
        driver.add_operator(op)
    driver.set_frequencies(freqs)
    driver.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    assert solver.solve_stats["lu_fallback"] == 0
    A = eqp.form_rpa_a_matrix_mo_singlet_partial(nv.to_dev(case["E"][0]), ovov, oovv)
    Bm = eqp.form_rpa_b_matrix_mo_singlet_partial(ovov)
    for f, w in enumerate(freqs):
        for op in ops:
            xy = op.rspvecs_alph[f][:, :, 0]              # (3, 2nov) on the device
            rhs = op.supervector_dev("alph")
            X, Y = xy[:, :nov], xy[:, nov:]
            top = X @ A.t() + Y @ Bm.t() - w * X
            bot = X @ Bm.t() + Y @ A.t() + w * Y
            resid = torch.cat((top, bot), dim=1) - rhs
            assert float(resid.abs().max()) <= 1e-11 * max(1.0, float(rhs.abs().max()))
        res = driver.results[f]
        assert res.shape == (9, 9)
        alpha = res[:3, :3]
        assert np.abs(alpha - alpha.T).max() <= 1e-10
        # linearity: the third operator is 2.5 x the first
        assert np.abs(res[6:, 6:] - 6.25 * alpha).max() <= 1e-9 * max(1.0, np.abs(alpha).max())
        assert np.abs(res[6:, :3] - 2.5 * alpha).max() <= 1e-9 * max(1.0, np.abs(alpha).max())


def test_solvers_share_hessian(golden_dir):
    """Two properties on one reference: the second solver adopts M, K, chol(K) of the first."""
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200.core import Hamiltonian, Spin

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o, build_eri=False)
    Or, Oi = synthetic.operator_integrals(n, 3, False), synthetic.operator_integrals(n, 3, True)
    tei = (g["ovov"], g["oovv"])
    s1 = solvers.ExactInvCholesky(case["C"], case["E"], case["occupations"])
    s1.tei_mo = tei
    d1 = cphf.CPHF(s1)
    d1.add_operator(operators.Operator(label="angmom", is_imaginary=True, ao_integrals=Oi))
    d1.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    s2 = solvers.ExactInv(case["C"], case["E"], case["occupations"])
    s2.tei_mo = tei
    s1.share_hessian_with(s2)
    d2 = cphf.CPHF(s2)
    d2.add_operator(operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or))
    d2.set_frequencies([float(w) for w in g["frequencies"]])
    d2.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    assert s2.solve_stats["potrf"] == 1          # factor of K came from the first solver
    ref = g["results_rpa_singlet_partial"]
    for f in range(len(g["frequencies"])):
        np.testing.assert_allclose(d2.results[f], ref[f][:3, :3], rtol=0, atol=1e-10)
    np.testing.assert_allclose(d1.results[0], ref[0][3:, 3:], rtol=0, atol=1e-10)


# ------------------------------------------------------------------------- round-2 API fixes
def test_lu_fallback_only_for_failed_frequencies(golden_dir):
    """Three frequencies, one above the first pole: the Cholesky results of the good ones are kept,
    LU runs once, M / K and the factor of K are not rebuilt."""
    from pymolresponse_b200 import synthetic

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o, build_eri=False)
    ops = [{"ao_integrals": synthetic.operator_integrals(n, 3, False), "is_imaginary": False},
           {"ao_integrals": synthetic.operator_integrals(n, 2, True), "is_imaginary": True}]
    freqs = [0.0, 1.9, 0.2]
    tei = (g["ovov"], g["oovv"])
    ref = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial", ops, freqs)
    solver, driver, objs = _run(case["C"], case["E"], case["occupations"], tei, "partial", ops, freqs,
                                "rpa", "singlet")
    assert solver.solve_stats["lu_fallback"] == 1
    assert solver.solve_stats["potrf"] == 1 and solver.solve_stats["potrf_batch"] == 3
    for f in range(3):
        np.testing.assert_allclose(driver.results[f], ref["results"][f], rtol=0, atol=1e-9)
        for k in range(2):
            assert np.abs(objs[k].rspvecs_alph[f] - ref["rspvecs_alph"][k][f]).max() <= 1e-9


def test_replaced_supervector_is_honoured(golden_dir):
    """The reference multiplies G^-1 by whatever ``mo_integrals_ai_supervector_alph`` holds when the
    solver runs (solvers.py:407-410): a caller that reassigns it after form_rhs (as
    form_rhs_geometric does) must get the response to the NEW, general [u; l] right-hand side."""
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200.core import Hamiltonian, Spin

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    nov = o * (n - o)
    case = synthetic.rhf_case(n, o, build_eri=False)
    tei = (g["ovov"], g["oovv"])
    w = 0.15
    solver = solvers.ExactInv(case["C"], case["E"], case["occupations"])
    solver.tei_mo = tei
    driver = cphf.CPHF(solver)
    op = operators.Operator(label="dipole", is_imaginary=False,
                            ao_integrals=synthetic.operator_integrals(n, 3, False))
    driver.add_operator(op)
    rng = np.random.default_rng(8)
    new_rhs = rng.standard_normal((3, 2 * nov, 1))
    op.mo_integrals_ai_supervector_alph = new_rhs
    driver.set_frequencies([w])
    driver.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    G = orc.explicit_hessian_rhf(case["E"], tei, "partial", o, "rpa", "singlet", w)
    want = np.linalg.solve(G, new_rhs[:, :, 0].T).T
    assert np.abs(op.rspvecs_alph[0][:, :, 0] - want).max() <= 1e-10
    np.testing.assert_allclose(driver.results[0], 2.0 * new_rhs[:, :, 0] @ want.T, rtol=0, atol=1e-9)


def test_cholesky_explicit_inverse_raises_and_uhf_left_blocks(golden_dir):
    from pymolresponse_b200 import solvers, synthetic

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o, build_eri=False)
    ops = [{"ao_integrals": synthetic.operator_integrals(n, 3, False), "is_imaginary": False}]
    tei = (g["ovov"], g["oovv"])
    s = solvers.ExactInvCholesky(case["C"], case["E"], case["occupations"])
    s.tei_mo = tei
    _, H, S = _enums()
    s.form_explicit_hessian(H["rpa"], S["singlet"], 0.1)
    s.invert_explicit_hessian()
    assert np.abs(s.explicit_hessian_inv @ s.explicit_hessian - np.eye(s.explicit_hessian.shape[0])).max() <= 1e-11
    s.form_explicit_hessian(H["rpa"], S["singlet"], 1.9)       # above the first pole: indefinite
    with pytest.raises(np.linalg.LinAlgError):
        s.invert_explicit_hessian()
    # UHF: left_alph / left_beta are readable right after run(), as in the reference
    gu = np.load(golden_dir / "synthetic_uhf_n9.npz")
    nu, na, nb = int(gu["n"]), int(gu["na"]), int(gu["nb"])
    ucase = synthetic.uhf_case(nu, na, nb, build_eri=False)
    uops = [{"ao_integrals": synthetic.operator_integrals(nu, 3, False), "is_imaginary": False}]
    utei = tuple(gu[f"partial_{k}"] for k in range(6))
    for cls in (solvers.ExactInv, solvers.ExactInvCholesky):
        su, _, _ = _run(ucase["C"], ucase["E"], ucase["occupations"], utei, "partial", uops, [0.1],
                        "rpa", "singlet", solver_cls=cls)
        G4 = orc.explicit_hessian_uhf(ucase["E"], utei, "partial", ucase["occupations"], "rpa",
                                      "singlet", 0.1)
        left_a = np.linalg.inv(G4[0] - G4[1] @ np.linalg.inv(G4[3]) @ G4[2])
        left_b = np.linalg.inv(G4[3] - G4[2] @ np.linalg.inv(G4[0]) @ G4[1])
        assert np.abs(su.left_alph - left_a).max() <= 1e-10
        assert np.abs(su.left_beta - left_b).max() <= 1e-10


# ------------------------------------------------------------------------- matrix-free block PCG
def test_iterative_cg_matches_reference_iterative_vectors(golden_dir):
    """method="cg": same response vectors as the reference's fixed-point IterativeLinEqSolver
    (golden vectors generated by the unmodified reference, solvers.py:603-637), from the MO blocks."""
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200.core import Hamiltonian, Spin

    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o, build_eri=False)
    it = solvers.IterativeLinEqSolver(case["C"], case["E"], case["occupations"], maxiter=60, conv=1e-26)
    assert it.method == "cg"
    it.tei_mo = (g["ovov"], g["oovv"])
    drv = cphf.CPHF(it)
    op = operators.Operator(label="dipole", is_imaginary=False,
                            ao_integrals=synthetic.operator_integrals(n, 3, False))
    drv.add_operator(op)
    drv.set_frequencies([0.0, 0.05])
    drv.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    assert it.iterations[-1] > 0
    np.testing.assert_allclose(np.stack(op.rspvecs_alph), g["iterative_rspvecs"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(np.stack(drv.results), g["iterative_results"], rtol=0, atol=1e-10)


@pytest.mark.parametrize("ham,spin", [("rpa", "singlet"), ("rpa", "triplet"), ("tda", "singlet")])
def test_iterative_cg_vs_exact_and_oracle_iterative(ham, spin):
    """nbasis = 64 (nov = 448): property tensors of the PCG path against ExactInv on the same MO
    blocks (<= 1e-8), response vectors against the exact ones at -w (the reference iterative solver's
    convention), real + imaginary operators, three frequencies in one block iteration, starting from
    the AO tensor (form_tei_mo)."""
    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200.core import Program

    _, H, S = _enums()
    n, o = 64, 8
    nov = o * (n - o)
    case = synthetic.rhf_case(n, o)
    Or, Oi = synthetic.operator_integrals(n, 3, False), synthetic.operator_integrals(n, 2, True)
    freqs = [0.0, 0.2, 0.5]
    it = solvers.IterativeLinEqSolver(case["C"], case["E"], case["occupations"], maxiter=80, conv=1e-26)
    drv = cphf.CPHF(it)
    ops = [operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or),
           operators.Operator(label="angmom", is_imaginary=True, ao_integrals=Oi)]
    for op in ops:
        drv.add_operator(op)
    drv.set_frequencies(freqs)
    drv.run(hamiltonian=H[ham], spin=S[spin], program=Program.Arrays, program_obj=case["I"])
    assert it.iterations[-1] > 0 and it.residuals[-1] < 1e-12
    tei = orc.ao2mo_rhf_partial(case["I"], case["C"], o)
    spec = [{"ao_integrals": Or, "is_imaginary": False}, {"ao_integrals": Oi, "is_imaginary": True}]
    ref = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial", spec, freqs, ham, spin)
    neg = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, "partial", spec,
                       [-w for w in freqs], ham, spin)
    for f in range(3):
        # the real-real block is independent of the +-w convention; for an
        # imaginary operator the reference's iterative algorithm returns the NEGATIVE of the exact
        # vectors (checked with the oracle port: 2 P.R^T = -results of ExactInv), mirrored here
        assert np.abs(drv.results[f][:3, :3] - ref["results"][f][:3, :3]).max() <= 1e-8
        assert np.abs(drv.results[f][3:, 3:] + ref["results"][f][3:, 3:]).max() <= 1e-8
        assert np.abs(ops[0].rspvecs_alph[f] - neg["rspvecs_alph"][0][f]).max() <= 1e-10
        assert np.abs(ops[1].rspvecs_alph[f] + neg["rspvecs_alph"][1][f]).max() <= 1e-10
        assert ops[0].rspvecs_alph[f].shape == (3, 2 * nov, 1)
    if ham == "rpa" and spin == "singlet":
        # and the reference's own fixed-point algorithm (oracle port, dense JK) on one operator
        vec, _ = orc.iterative_rhf(case["C"], case["E"], case["occupations"], Or, 0.2,
                                   orc.DenseJK(case["I"]), maxiter=200, conv=1e-26)
        assert np.abs(ops[0].rspvecs_alph[1] - vec).max() <= 1e-10


def test_row_block_assembly_matches_full_matrices():
    """pmr_assemble's [i_lo, i_hi) row range with row-sharded inputs (only the occupied rows a rank
    owns): every row block is bit-identical to the same rows of the full A / B / M / K, for ragged
    splits, the fast (even nvirt) and the generic (odd nvirt) kernels, ss / os kinds."""
    from pymolresponse_b200 import _blocks as bl
    from pymolresponse_b200 import _native as nv

    r = np.random.default_rng(21)
    for o, v in ((5, 12), (4, 7)):
        nov = o * v
        ovov = r.standard_normal((o, v, o, v))
        oovv = r.standard_normal((o, o, v, v))
        E = np.diag(np.sort(r.standard_normal(o + v)))
        eps = bl.split_energies(E, o)
        full_A = orc.a_singlet_partial(E, ovov, oovv)
        full_B = orc.b_singlet_partial(ovov)
        for lo, hi in ((0, 2), (2, 3), (3, o)):
            t0, t1 = nv.to_dev(ovov[lo:hi]), nv.to_dev(oovv[lo:hi])
            terms = bl.partial_terms("singlet", t0, t1, lo)
            A, B = bl.run(o, v, o, v, terms, eps, "PQ", i_lo=lo, i_hi=hi)
            assert tuple(A.shape) == ((hi - lo) * v, nov)
            assert np.array_equal(nv.to_host(A), full_A[lo * v:hi * v])
            assert np.array_equal(nv.to_host(B), full_B[lo * v:hi * v])
            M, K = nv.empty((hi - lo) * v, nov), nv.empty((hi - lo) * v, nov)
            bl.run(o, v, o, v, terms, eps, "PQ", out_mode=1, out0=M, out1=K, ld=nov, i_lo=lo, i_hi=hi)
            assert np.array_equal(nv.to_host(M), (full_A - full_B)[lo * v:hi * v])
            assert np.array_equal(nv.to_host(K), (full_A + full_B)[lo * v:hi * v])
        # rectangular opposite-spin block, rows sharded
        xxyy = r.standard_normal((o, v, 3, v + 1))
        full_os = orc.a_singlet_os_partial(xxyy)
        lo, hi = 1, 4
        terms = bl.partial_terms("os", nv.to_dev(xxyy[lo:hi]), None, lo)
        A_os, B_os = bl.run(o, v, 3, v + 1, terms, None, "PQ", i_lo=lo, i_hi=hi)
        assert np.array_equal(nv.to_host(A_os), full_os[lo * v:hi * v])
        assert np.array_equal(nv.to_host(B_os), -full_os[lo * v:hi * v])


def test_form_tei_mo_through_pyscf_and_psi4_program_objects(monkeypatch):
    """Solver.form_tei_mo (reference solvers.py:86-127) with Program.PySCF / Program.Psi4 objects:
    duck-typed stand-ins hand out the AO ERI tensor exactly as pyscf's ``mol.intor('int2e',
    aosym='s1')`` / psi4's ``MintsHelper.ao_eri()`` would; results equal the Program.Arrays route."""
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from _fake_qc import FakeMole, install_fake_psi4

    from pymolresponse_b200 import cphf, operators, solvers, synthetic
    from pymolresponse_b200.core import Hamiltonian, Program, Spin

    n, o = 20, 3
    case = synthetic.rhf_case(n, o)
    Or = synthetic.operator_integrals(n, 3, False)
    install_fake_psi4(monkeypatch, n, eri=case["I"])
    results = {}
    for tag, program, obj in (("arrays", Program.Arrays, case["I"]),
                              ("pyscf", Program.PySCF, FakeMole(n, (o, o), eri=case["I"])),
                              ("psi4", Program.Psi4, "fake psi4 molecule")):
        s = solvers.ExactInv(case["C"], case["E"], case["occupations"])
        d = cphf.CPHF(s)
        d.add_operator(operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or))
        d.set_frequencies([0.0, 0.1])
        d.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=program, program_obj=obj)
        results[tag] = np.stack(d.results)
    assert np.array_equal(results["pyscf"], results["arrays"])
    assert np.array_equal(results["psi4"], results["arrays"])
