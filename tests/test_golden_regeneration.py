"""CPU, build container only: the committed golden vectors are what the UNMODIFIED reference produces
today.  Re-runs pieces of tests/golden/make_golden.py against /root/reference and compares with the
stored .npz files; skipped where the reference is not mounted (the GPU box)."""

import importlib.util
import os
import sys
from pathlib import Path

import numpy as np
import pytest

REFERENCE = Path("/root/reference/pymolresponse")
pytestmark = pytest.mark.skipif(not REFERENCE.exists(), reason="reference not mounted here")


@pytest.fixture(scope="module")
def mg():
    path = Path(__file__).resolve().parent / "golden" / "make_golden.py"
    spec = importlib.util.spec_from_file_location("make_golden_for_tests", path)
    module = importlib.util.module_from_spec(spec)
    saved = list(sys.path)
    try:
        spec.loader.exec_module(module)
    finally:
        # the script puts /root/reference on sys.path; the other tests must not see it
        sys.path[:] = saved
    return module


def test_water_results_regenerate(mg, golden_dir):
    from pymolresponse import utils
    from pymolresponse.core import AO2MOTransformationType
    from pymolresponse.data import REFDIR

    g = np.load(golden_dir / "water_sto3g.npz")
    C = utils.fix_mocoeffs_shape(utils.np_load_2(REFDIR / "C.npz"))
    E = utils.fix_moenergies_shape(utils.np_load_2(REFDIR / "F_MO.npz"))
    TEI = utils.np_load(REFDIR / "TEI_MO.npz")
    assert np.array_equal(C, g["C"]) and np.array_equal(TEI, g["TEI_MO"])
    for ham, spin in (("rpa", "singlet"), ("tda", "triplet")):
        _, driver, ops = mg.run_reference(
            C, E, (5, 2, 5, 2), (TEI,), AO2MOTransformationType.full,
            [{"is_imaginary": False, "ao_integrals": g["dipole"]}], g["frequencies"], ham, spin)
        assert np.array_equal(np.stack(driver.results), g[f"results_{ham}_{spin}_inv"])
        assert np.array_equal(np.stack(ops[0].rspvecs_alph), g[f"rspvecs_{ham}_{spin}_inv"])


def test_excitation_energies_regenerate(mg, golden_dir):
    import contextlib
    import io

    from pymolresponse.core import AO2MOTransformationType

    g = np.load(golden_dir / "td_golden.npz")
    w = np.load(golden_dir / "water_sto3g.npz")
    ops = [{"label": "angmom", "is_imaginary": True, "ao_integrals": g["water_angmom_like"]},
           {"label": "dipole", "is_imaginary": False, "ao_integrals": w["dipole"]}]
    with contextlib.redirect_stdout(io.StringIO()):
        solver, op_objs, ecd = mg._run_td(w["C"], w["E"], (5, 2, 5, 2), (w["TEI_MO"],),
                                          AO2MOTransformationType.full, ops, "rpa", "singlet", False)
    np.testing.assert_allclose(solver.eigvals.real, g["eigvals_water_rpasolver_rpa_singlet"], rtol=0,
                               atol=1e-13)
    np.testing.assert_allclose(op_objs[1].total_oscillator_strengths,
                               g["totosc_dipole_water_rpasolver_rpa_singlet"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(ecd.rotational_strengths_diplen,
                               g["rotstr_water_rpasolver_rpa_singlet"], rtol=1e-10, atol=1e-10)


def test_small_helpers_agree_with_the_reference(mg):
    """utils helpers that have no numerics of their own on the GPU: same answers as the reference's."""
    from pymolresponse import utils as ref_utils

    from pymolresponse_b200 import utils

    rng = np.random.default_rng(11)
    for trial in range(6):
        A = rng.standard_normal((6, 6))
        mats = [A, A + A.T, A - A.T, np.zeros((4, 4)), 1e-15 * A]
        for M in mats:
            assert utils.matsym(M) == ref_utils.matsym(M)
            for tri in ("lower", "upper"):
                assert np.array_equal(utils.flip_triangle_sign(M, tri), ref_utils.flip_triangle_sign(M, tri))
            assert np.array_equal(utils.screen(M, 0.3), ref_utils.screen(M, 0.3))
    widths = (4, 3, 5, 6, 10, 10)
    line = "   1  C 1   s       -0.00000  -0.0"
    for trunc in (True, False):
        assert utils.Splitter(widths).split(line, trunc) == ref_utils.Splitter(widths).split(line, trunc)
    assert utils.form_indices_orbwin(3, 4) == ref_utils.form_indices_orbwin(3, 4)
    assert utils.form_indices_zero(3, 4) == ref_utils.form_indices_zero(3, 4)
    moene_occ, moene_virt = rng.standard_normal(3), rng.standard_normal(4)
    want = ref_utils.form_vec_energy_differences(moene_occ, moene_virt)
    # the GPU version is covered by the -m gpu tests; here only the layout contract of the oracle
    from oracle import pmr_oracle as orc

    assert np.array_equal(orc.form_vec_energy_differences(moene_occ, moene_virt), want)


@pytest.mark.parametrize("n,nocc", [(40, 5), (64, 8)])
def test_oracle_equals_live_reference_at_larger_sizes(mg, n, nocc):
    """Beyond the stored n = 10 vectors: the oracle against the imported reference on a seeded
    synthetic RHF case with nov in the hundreds (AO2MO, A/B blocks, ExactInv at two frequencies,
    mixed real / imaginary operators)."""
    from pymolresponse import explicit_equations_partial as eqp
    from pymolresponse.ao2mo import AO2MO
    from pymolresponse.core import AO2MOTransformationType

    from oracle import pmr_oracle as orc
    from pymolresponse_b200 import synthetic

    case = synthetic.rhf_case(n, nocc)
    C, E, occ, I = case["C"], case["E"], case["occupations"], case["I"]
    a = AO2MO(C, occ, I=I)
    a.perform_rhf_partial()
    ovov, oovv = a.tei_mo
    o_ovov, o_oovv = orc.ao2mo_rhf_partial(I, C, nocc)
    scale = np.abs(ovov).max()
    assert np.abs(o_ovov - ovov).max() <= 1e-12 * scale
    assert np.abs(o_oovv - oovv).max() <= 1e-12 * np.abs(oovv).max()
    assert np.array_equal(orc.a_singlet_partial(E[0], ovov, oovv),
                          eqp.form_rpa_a_matrix_mo_singlet_partial(E[0], ovov, oovv))
    assert np.array_equal(orc.b_triplet_partial(ovov), eqp.form_rpa_b_matrix_mo_triplet_partial(ovov))
    ops = [{"label": "dipole", "is_imaginary": False,
            "ao_integrals": synthetic.operator_integrals(n, 3, False)},
           {"label": "angmom", "is_imaginary": True,
            "ao_integrals": synthetic.operator_integrals(n, 3, True)}]
    freqs = [0.0, 0.11]
    _, driver, op_objs = mg.run_reference(C, E, occ, (ovov, oovv), AO2MOTransformationType.partial,
                                          ops, freqs, "rpa", "singlet")
    out = orc.run_cphf(C, E, occ, (ovov, oovv), "partial", ops, freqs, "rpa", "singlet")
    for f in range(2):
        np.testing.assert_allclose(out["results"][f], driver.results[f], rtol=0, atol=1e-11)
        np.testing.assert_allclose(out["uncoupled_results"][f], driver.uncoupled_results[f], rtol=0,
                                   atol=1e-12)
        np.testing.assert_allclose(out["rspvecs_alph"][0][f], op_objs[0].rspvecs_alph[f], rtol=0,
                                   atol=1e-11)


def test_oracle_equals_live_reference_uhf(mg):
    """UHF (7 alpha, 5 beta electrons, n = 30): six partial MO blocks from the reference's own
    transform, joint solve with block elimination (solvers.py:426-466) at two frequencies."""
    from pymolresponse.ao2mo import AO2MO
    from pymolresponse.core import AO2MOTransformationType

    from oracle import pmr_oracle as orc
    from pymolresponse_b200 import synthetic

    n, na, nb = 30, 7, 5
    case = synthetic.uhf_case(n, na, nb)
    C, E, occ, I = case["C"], case["E"], case["occupations"], case["I"]
    Coa, Cva, Cob, Cvb = C[0][:, :na], C[0][:, na:], C[1][:, :nb], C[1][:, nb:]
    T = AO2MO.transform
    partial = (T(I, Coa, Cva, Coa, Cva), T(I, Coa, Cva, Cob, Cvb), T(I, Cob, Cvb, Coa, Cva),
               T(I, Cob, Cvb, Cob, Cvb), T(I, Coa, Coa, Cva, Cva), T(I, Cob, Cob, Cvb, Cvb))
    mine = orc.ao2mo_uhf_partial(I, C, occ)
    for got, want in zip(mine, partial):
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
    ops = [{"label": "dipole", "is_imaginary": False,
            "ao_integrals": synthetic.operator_integrals(n, 3, False)},
           {"label": "angmom", "is_imaginary": True,
            "ao_integrals": synthetic.operator_integrals(n, 3, True)}]
    freqs = [0.0, 0.09]
    for ham, spin in (("rpa", "singlet"), ("tda", "triplet")):
        _, driver, op_objs = mg.run_reference(C, E, occ, partial, AO2MOTransformationType.partial, ops,
                                              freqs, ham, spin)
        out = orc.run_cphf(C, E, occ, partial, "partial", ops, freqs, ham, spin)
        for f in range(2):
            np.testing.assert_allclose(out["results"][f], driver.results[f], rtol=0, atol=1e-11)
            np.testing.assert_allclose(out["rspvecs_beta"][1][f], op_objs[1].rspvecs_beta[f], rtol=0,
                                       atol=1e-11)


def _factorised_vs_live_reference(mg, n, nocc, freqs):
    from pymolresponse.ao2mo import AO2MO
    from pymolresponse.core import AO2MOTransformationType

    from oracle import pmr_factorised as fac
    from pymolresponse_b200 import synthetic

    case = synthetic.rhf_case(n, nocc)
    C, E, occ, I = case["C"], case["E"], case["occupations"], case["I"]
    a = AO2MO(C, occ, I=I)
    a.perform_rhf_partial()
    ovov, oovv = a.tei_mo
    del I, a
    f_ovov, f_oovv = fac.mo_blocks_rhf(case["B"], C[0], nocc)
    assert np.abs(f_ovov - ovov).max() <= 1e-12 * np.abs(ovov).max()
    assert np.abs(f_oovv - oovv).max() <= 1e-12 * np.abs(oovv).max()
    ops = [{"label": "dipole", "is_imaginary": False,
            "ao_integrals": synthetic.operator_integrals(n, 3, False)},
           {"label": "angmom", "is_imaginary": True,
            "ao_integrals": synthetic.operator_integrals(n, 3, True)}]
    _, driver, op_objs = mg.run_reference(C, E, occ, (ovov, oovv), AO2MOTransformationType.partial,
                                          ops, freqs, "rpa", "singlet")
    got = fac.run_cphf_factorised(case["B"], C, E, occ, ops, freqs)
    worst = 0.0
    for f in range(len(freqs)):
        worst = max(worst, float(np.abs(got["results"][f] - driver.results[f]).max()))
        np.testing.assert_allclose(got["results"][f], driver.results[f], rtol=0, atol=1e-10)
        for k, (c0, c1) in enumerate(got["spans"]):
            np.testing.assert_allclose(got["xy_alph"][f, c0:c1], op_objs[k].rspvecs_alph[f][:, :, 0],
                                       rtol=0, atol=1e-10)
    return worst


def test_factorised_restatement_equals_live_reference_n96(mg):
    """oracle/pmr_factorised.py (the checker used at n = 256 / 384 / 512) against the unmodified
    reference: AO2MO.perform_rhf_partial on the dense tensor, ExactInv at a static and two dynamic
    frequencies, real + imaginary operators."""
    _factorised_vs_live_reference(mg, 96, 12, [0.0, 0.2, 0.55])


@pytest.mark.skipif(os.environ.get("PMR_SLOW", "0") != "1", reason="minutes of CPU: set PMR_SLOW=1")
def test_factorised_restatement_equals_live_reference_n160(mg):
    """Largest size the reference finishes here in minutes (I = 5.2 GB, 2 nov = 5600)."""
    worst = _factorised_vs_live_reference(mg, 160, 20, [0.0, 0.35])
    print(f"n=160: max |d results| vs live reference = {worst:.3e}")


def test_factorised_restatement_equals_live_reference_uhf(mg):
    from pymolresponse.ao2mo import AO2MO
    from pymolresponse.core import AO2MOTransformationType

    from oracle import pmr_factorised as fac
    from pymolresponse_b200 import synthetic

    n, na, nb = 40, 6, 5
    case = synthetic.uhf_case(n, na, nb)
    C, E, occ, I = case["C"], case["E"], case["occupations"], case["I"]
    Coa, Cva, Cob, Cvb = C[0][:, :na], C[0][:, na:], C[1][:, :nb], C[1][:, nb:]
    T = AO2MO.transform
    partial = (T(I, Coa, Cva, Coa, Cva), T(I, Coa, Cva, Cob, Cvb), T(I, Cob, Cvb, Coa, Cva),
               T(I, Cob, Cvb, Cob, Cvb), T(I, Coa, Coa, Cva, Cva), T(I, Cob, Cob, Cvb, Cvb))
    for got, want in zip(fac.mo_blocks_uhf(case["B"], C, occ), partial):
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    ops = [{"label": "dipole", "is_imaginary": False,
            "ao_integrals": synthetic.operator_integrals(n, 3, False)}]
    freqs = [0.0, 0.12]
    _, driver, op_objs = mg.run_reference(C, E, occ, partial, AO2MOTransformationType.partial, ops,
                                          freqs, "rpa", "singlet")
    got = fac.run_cphf_factorised(case["B"], C, E, occ, ops, freqs)
    for f in range(2):
        np.testing.assert_allclose(got["results"][f], driver.results[f], rtol=0, atol=1e-10)
        np.testing.assert_allclose(got["xy_alph"][f], op_objs[0].rspvecs_alph[f][:, :, 0], rtol=0, atol=1e-10)
        np.testing.assert_allclose(got["xy_beta"][f], op_objs[0].rspvecs_beta[f][:, :, 0], rtol=0, atol=1e-10)
