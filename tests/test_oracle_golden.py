"""CPU: pin the NumPy oracle (oracle/pmr_oracle.py) against
  (1) the literals of the reference's own known-answer tests (water/STO-3G),
  (2) outputs of the unmodified reference stored by tests/golden/make_golden.py,
  (3) the DALTON values of the LiH / LiH+ disk fixtures with the reference's thresholds.
"""

import numpy as np
import pytest

from oracle import pmr_oracle as orc

FULL, PARTIAL = "full", "partial"

# literals of /root/reference/pymolresponse/tests/test_final_result.py:54-63, :135-146, :216-225,
# :295-306 (diagonals of the 3x3 polarizability tensors at w = 0, 0.02, 0.06, 0.1; off-diagonals 0)
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


def _water(golden_dir):
    return np.load(golden_dir / "water_sto3g.npz")


@pytest.mark.parametrize("ham", ["rpa", "tda"])
@pytest.mark.parametrize("spin", ["singlet", "triplet"])
@pytest.mark.parametrize("chol", [False, True])
def test_oracle_water_fixture(golden_dir, ham, spin, chol):
    g = _water(golden_dir)
    out = orc.run_cphf(g["C"], g["E"], (5, 2, 5, 2), (g["TEI_MO"],), FULL,
                       [{"ao_integrals": g["dipole"], "is_imaginary": False}], g["frequencies"],
                       ham, spin, cholesky=chol)
    tag = f"{ham}_{spin}_{'chol' if chol else 'inv'}"
    np.testing.assert_allclose(np.stack(out["results"]), g[f"results_{tag}"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(np.stack(out["uncoupled_results"]), g[f"uncoupled_{tag}"], rtol=0,
                               atol=1e-12)
    np.testing.assert_allclose(np.stack(out["rspvecs_alph"][0]), g[f"rspvecs_{tag}"], rtol=0,
                               atol=1e-11)
    if (ham, spin) in WATER_LITERALS:
        for f, diag in enumerate(WATER_LITERALS[(ham, spin)]):
            np.testing.assert_allclose(out["results"][f], np.diag(diag), rtol=0, atol=1e-8)


def _rhf(golden_dir):
    return np.load(golden_dir / "synthetic_rhf_n10.npz")


def test_oracle_ao2mo_and_blocks_rhf(golden_dir):
    from pymolresponse_b200 import synthetic

    g = _rhf(golden_dir)
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o)
    ovov, oovv = orc.ao2mo_rhf_partial(case["I"], case["C"], o)
    (full,) = orc.ao2mo_rhf_full(case["I"], case["C"])
    for got, key in ((ovov, "ovov"), (oovv, "oovv"), (full, "full")):
        np.testing.assert_allclose(got, g[key], rtol=0, atol=1e-15)
    E = case["E"][0]
    checks = {
        "A_singlet_partial": orc.a_singlet_partial(E, ovov, oovv),
        "A_triplet_partial": orc.a_triplet_partial(E, oovv),
        "B_singlet_partial": orc.b_singlet_partial(ovov),
        "B_triplet_partial": orc.b_triplet_partial(ovov),
        "A_singlet_ss_partial": orc.a_singlet_ss_partial(E, ovov, oovv),
        "A_singlet_os_partial": orc.a_singlet_os_partial(ovov),
        "B_singlet_ss_partial": orc.b_singlet_ss_partial(ovov),
        "B_singlet_os_partial": orc.b_singlet_os_partial(ovov),
        "A_singlet_full": orc.a_singlet_full(E, full, o),
        "A_triplet_full": orc.a_triplet_full(E, full, o),
        "B_singlet_full": orc.b_singlet_full(full, o),
        "B_triplet_full": orc.b_triplet_full(full, o),
        "A_singlet_ss_full": orc.a_singlet_ss_full(E, full, o),
        "A_singlet_os_full": orc.a_singlet_os_full(full, o, o),
        "B_singlet_ss_full": orc.b_singlet_ss_full(full, o),
        "B_singlet_os_full": orc.b_singlet_os_full(full, o, o),
    }
    for key, got in checks.items():
        assert np.array_equal(got, g[key]), key  # same operation order => bit-identical


@pytest.mark.parametrize("ham", ["rpa", "tda"])
@pytest.mark.parametrize("spin", ["singlet", "triplet"])
@pytest.mark.parametrize("ttype", [PARTIAL, FULL])
def test_oracle_cphf_rhf(golden_dir, ham, spin, ttype):
    from pymolresponse_b200 import synthetic

    g = _rhf(golden_dir)
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o, build_eri=False)
    ops = [{"ao_integrals": synthetic.operator_integrals(n, 3, False), "is_imaginary": False},
           {"ao_integrals": synthetic.operator_integrals(n, 3, True), "is_imaginary": True}]
    tei = (g["ovov"], g["oovv"]) if ttype == PARTIAL else (g["full"],)
    out = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, ttype, ops, g["frequencies"],
                       ham, spin)
    tag = f"{ham}_{spin}_{ttype}"
    np.testing.assert_allclose(np.stack(out["results"]), g[f"results_{tag}"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.stack(out["uncoupled_results"]), g[f"uncoupled_{tag}"], rtol=0,
                               atol=1e-13)
    np.testing.assert_allclose(np.stack(out["rspvecs_alph"][0]), g[f"rspvecs0_{tag}"], rtol=0,
                               atol=1e-12)
    np.testing.assert_allclose(np.stack(out["rspvecs_alph"][1]), g[f"rspvecs1_{tag}"], rtol=0,
                               atol=1e-12)
    if ttype == PARTIAL:
        G = orc.explicit_hessian_rhf(case["E"], tei, ttype, o, ham, spin, float(g["frequencies"][-1]))
        assert np.array_equal(G, g[f"hessian_last_{tag}"])


def test_oracle_rhs_and_cholesky_rhf(golden_dir):
    from pymolresponse_b200 import synthetic

    g = _rhf(golden_dir)
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o, build_eri=False)
    Or, Oi = synthetic.operator_integrals(n, 3, False), synthetic.operator_integrals(n, 3, True)
    r0 = orc.form_rhs(Or, case["C"], case["occupations"], is_imaginary=False)
    r1 = orc.form_rhs(Oi, case["C"], case["occupations"], is_imaginary=True)
    assert np.array_equal(r0["mo_integrals_ai_alph"], g["rhs0_ai"])
    assert np.array_equal(r0["mo_integrals_ai_supervector_alph"], g["rhs0_super"])
    assert np.array_equal(r1["mo_integrals_ai_supervector_alph"], g["rhs1_super"])
    ops = [{"ao_integrals": Or, "is_imaginary": False}, {"ao_integrals": Oi, "is_imaginary": True}]
    out = orc.run_cphf(case["C"], case["E"], case["occupations"], (g["ovov"], g["oovv"]), PARTIAL,
                       ops, g["frequencies"], "rpa", "singlet", cholesky=True)
    np.testing.assert_allclose(np.stack(out["results"]), g["results_rpa_singlet_partial_chol"],
                               rtol=0, atol=1e-12)


def test_oracle_iterative_rhf(golden_dir):
    from pymolresponse_b200 import synthetic

    g = _rhf(golden_dir)
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o)
    Or = synthetic.operator_integrals(n, 3, False)
    jk = orc.DenseJK(case["I"])
    for f, w in enumerate(g["iterative_frequencies"]):
        vec, it = orc.iterative_rhf(case["C"], case["E"], case["occupations"], Or, float(w), jk,
                                    maxiter=200, conv=1e-24)
        assert vec is not None
        np.testing.assert_allclose(vec, g["iterative_rspvecs"][f], rtol=0, atol=1e-13)
        sv = orc.form_rhs(Or, case["C"], case["occupations"])["mo_integrals_ai_supervector_alph"]
        res = orc.coupled_results([sv], [vec])
        np.testing.assert_allclose(res, g["iterative_results"][f], rtol=0, atol=1e-13)
        # the iterative property tensor agrees with the exact one (tests/test_solvers.py:108-113)
        np.testing.assert_allclose(res, g["results_rpa_singlet_partial"][f][:3, :3], rtol=0, atol=1e-8)


def _uhf(golden_dir):
    return np.load(golden_dir / "synthetic_uhf_n9.npz")


def test_oracle_ao2mo_uhf(golden_dir):
    from pymolresponse_b200 import synthetic

    g = _uhf(golden_dir)
    n, na, nb = int(g["n"]), int(g["na"]), int(g["nb"])
    case = synthetic.uhf_case(n, na, nb)
    part = orc.ao2mo_uhf_partial(case["I"], case["C"], case["occupations"])
    full = orc.ao2mo_uhf_full(case["I"], case["C"])
    for k, t in enumerate(part):
        np.testing.assert_allclose(t, g[f"partial_{k}"], rtol=0, atol=1e-15)
    for k, t in enumerate(full):
        np.testing.assert_allclose(t, g[f"full_{k}"], rtol=0, atol=1e-15)
    assert np.array_equal(orc.a_singlet_os_full(full[1], na, nb), g["A_os_a_full"])
    assert np.array_equal(orc.b_singlet_os_full(full[2], nb, na), g["B_os_b_full"])


@pytest.mark.parametrize("ham", ["rpa", "tda"])
@pytest.mark.parametrize("spin", ["singlet", "triplet"])
@pytest.mark.parametrize("ttype", [PARTIAL, FULL])
def test_oracle_cphf_uhf(golden_dir, ham, spin, ttype):
    from pymolresponse_b200 import synthetic

    g = _uhf(golden_dir)
    n, na, nb = int(g["n"]), int(g["na"]), int(g["nb"])
    case = synthetic.uhf_case(n, na, nb, build_eri=False)
    ops = [{"ao_integrals": synthetic.operator_integrals(n, 3, False), "is_imaginary": False},
           {"ao_integrals": synthetic.operator_integrals(n, 3, True), "is_imaginary": True}]
    tei = tuple(g[f"{ttype}_{k}"] for k in range(6 if ttype == PARTIAL else 4))
    out = orc.run_cphf(case["C"], case["E"], case["occupations"], tei, ttype, ops, g["frequencies"],
                       ham, spin)
    tag = f"{ham}_{spin}_{ttype}"
    np.testing.assert_allclose(np.stack(out["results"]), g[f"results_{tag}"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(np.stack(out["uncoupled_results"]), g[f"uncoupled_{tag}"], rtol=0,
                               atol=1e-13)
    for k in (0, 1):
        np.testing.assert_allclose(np.stack(out["rspvecs_alph"][k]), g[f"rspvecs{k}_alph_{tag}"],
                                   rtol=0, atol=1e-12)
        np.testing.assert_allclose(np.stack(out["rspvecs_beta"][k]), g[f"rspvecs{k}_beta_{tag}"],
                                   rtol=0, atol=1e-12)


def test_oracle_uhf_hessian_blocks(golden_dir):
    from pymolresponse_b200 import synthetic

    g = _uhf(golden_dir)
    n, na, nb = int(g["n"]), int(g["na"]), int(g["nb"])
    case = synthetic.uhf_case(n, na, nb, build_eri=False)
    tei = tuple(g[f"partial_{k}"] for k in range(6))
    G4 = orc.explicit_hessian_uhf(case["E"], tei, PARTIAL, case["occupations"], "rpa", "singlet",
                                  float(g["frequencies"][-1]))
    for q in range(4):
        assert np.array_equal(G4[q], g[f"hessian_last_{q}"])


@pytest.mark.parametrize("case,uhf,thresh", [("r_lih_hf_sto-3g", False, 5.0e-3),
                                             ("u_lih_cation_hf_sto-3g", True, 1.0e-1)])
def test_oracle_lih_dalton_fixtures(golden_dir, case, uhf, thresh):
    """tests/test_runners.py:27-100: every DALTON value of the disk fixtures (same signed
    threshold test as the reference), and the reference's own numbers to 1e-10.  A strided sample
    keeps the CPU suite short; the GPU suite runs another sample through the CUDA path."""
    g = np.load(golden_dir / f"{case}.npz")
    occ = tuple(int(x) for x in g["occupations"])
    if uhf:
        E = np.stack((np.diag(g["moene"][:, 0]), np.diag(g["moene"][:, 1])), axis=0)
        tei = tuple(g[k] for k in ("moints_iajb_aaaa", "moints_iajb_aabb", "moints_iajb_bbaa",
                                   "moints_iajb_bbbb", "moints_ijab_aaaa", "moints_ijab_bbbb"))
    else:
        E = np.diag(g["moene"][:, 0])[None]
        tei = (g["moints_iajb_aaaa"], g["moints_ijab_aaaa"])
    names = [str(x) for x in g["label_names"]]
    idx = [k for k in range(len(g["ref_value"])) if not np.isnan(g["reference_value"][k])]
    checked = 0
    for k in idx[::7]:
        ops = []
        for lab in (str(g["ref_label_1"][k]), str(g["ref_label_2"][k])):
            q = names.index(lab)
            ints = g[f"operator_mn_{g['label_operator'][q]}"][int(g["label_slice"][q])]
            op = {"ao_integrals": ints[None], "is_imaginary": bool(g["label_imaginary"][q])}
            if "spinorb" in str(g["label_operator"][q]):
                op["hsofac"] = (0.0072973525693 ** 2) / 4
            ops.append(op)
        out = orc.run_cphf(g["C"], E, occ, tei, PARTIAL, ops, [float(g["ref_frequency"][k])],
                           str(g["ref_hamiltonian"][k]), str(g["ref_spin"][k]))
        res = out["results"][0]
        assert abs(abs(res[1, 0]) - abs(res[0, 1])) < 1e-13
        assert abs(res[1, 0] - g["reference_value"][k]) < 1e-10
        assert abs(g["ref_value"][k]) - abs(res[1, 0]) < thresh
        checked += 1
    assert checked > 500


# ------------------------------------------------------------------ excitation energies
@pytest.mark.parametrize("name", ["water", "syn10", "lih"])
@pytest.mark.parametrize("tda_solver,ham", [(False, "rpa"), (False, "tda"), (True, "tda")])
@pytest.mark.parametrize("spin", ["singlet", "triplet"])
def test_oracle_excitation_energies(golden_dir, name, tda_solver, ham, spin):
    """diagonalize_rhf / transition_results against the reference's ExactDiagonalizationSolver(TDA)
    + td.TDHF / td.TDA + ECD.form_results (signs of eigenvectors are arbitrary: moments compared
    in magnitude, strengths as they are)."""
    from _td_cases import tag_of, td_system

    from pymolresponse_b200.constants import esuecd

    sysd = td_system(golden_dir, name)
    g = np.load(golden_dir / "td_golden.npz")
    tag = tag_of(name, tda_solver, ham, spin)
    nocc, nvirt = sysd["occ"][0], sysd["occ"][1]
    ev, vecs = orc.diagonalize_rhf(sysd["E"], sysd["tei"], sysd["ttype"], nocc, ham, spin,
                                   tda_solver=tda_solver)
    np.testing.assert_allclose(ev.real, g[f"eigvals_{tag}"], rtol=0, atol=1e-12)
    key = "mo_integrals_ai_alph" if tda_solver else "mo_integrals_ai_supervector_alph"
    grads = [orc.form_rhs(sysd["Oimag"], sysd["C"], sysd["occ"], is_imaginary=True)[key][:, :, 0],
             orc.form_rhs(sysd["Oreal"], sysd["C"], sysd["occ"], is_imaginary=False)[key][:, :, 0]]
    res = orc.transition_results(ev, vecs, grads, nocc, nvirt, tda_driver=tda_solver)
    for label, r in zip(("angmom", "dipole"), res):
        np.testing.assert_allclose(np.abs(r["transition_moments"]), np.abs(g[f"tmom_{label}_{tag}"]),
                                   rtol=0, atol=1e-10)
        np.testing.assert_allclose(r["oscillator_strengths"], g[f"osc_{label}_{tag}"], rtol=0,
                                   atol=1e-10)
        np.testing.assert_allclose(r["total_oscillator_strengths"], g[f"totosc_{label}_{tag}"],
                                   rtol=0, atol=1e-10)
    rot = esuecd * (-1 / 2) * np.einsum("sc,sc->s", res[1]["transition_moments"],
                                        res[0]["transition_moments"])
    np.testing.assert_allclose(rot, g[f"rotstr_{tag}"], rtol=0, atol=1e-7 * max(1.0, np.abs(rot).max()))
