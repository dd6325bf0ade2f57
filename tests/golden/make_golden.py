"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` is mounted there and absent on the GPU
box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference is imported from ``/root/reference`` with a stub ``pymolresponse._version``
module (``pymolresponse/__init__.py:3`` imports a versioningit-generated file that a source
checkout does not have).  Nothing from the reference's sources is copied: only its inputs
(on-disk data fixtures) and its numerical outputs are stored, as ``.npz``.
"""

from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
sys.path.insert(0, str(REPO))

_v = types.ModuleType("pymolresponse._version")
_v.__version__ = "0+oracle"
sys.modules["pymolresponse._version"] = _v
# ``pymolresponse/constants.py:3,6,22`` builds a pint unit registry for one dipole conversion
# constant; pint is not installed here and none of the numbers stored below depend on it.
_pint = types.ModuleType("pint")


class _UnitRegistry:
    debye = None

    def __init__(self, **_kw):
        pass

    def get_base_units(self, input_units=None):
        return (1.0, None)


_pint.UnitRegistry = _UnitRegistry
sys.modules.setdefault("pint", _pint)
sys.path.insert(0, "/root/reference")

from pymolresponse import cphf, operators, solvers, utils  # noqa: E402
from pymolresponse import explicit_equations_full as eqf  # noqa: E402
from pymolresponse import explicit_equations_partial as eqp  # noqa: E402
from pymolresponse.ao2mo import AO2MO  # noqa: E402
from pymolresponse.core import AO2MOTransformationType, Hamiltonian, Spin  # noqa: E402
from pymolresponse.data import REFDIR  # noqa: E402
from pymolresponse.integrals import JK  # noqa: E402
from pymolresponse.interfaces.dalton.utils import dalton_label_to_operator  # noqa: E402

from pymolresponse_b200 import synthetic  # noqa: E402

HAMS = {"rpa": Hamiltonian.RPA, "tda": Hamiltonian.TDA}
SPINS = {"singlet": Spin.singlet, "triplet": Spin.triplet}


def run_reference(C, E, occ, tei_mo, tei_type, ops, freqs, ham, spin, cholesky=False):
    cls = solvers.ExactInvCholesky if cholesky else solvers.ExactInv
    solver = cls(C, E, occ)
    solver.tei_mo = tei_mo
    solver.tei_mo_type = tei_type
    driver = cphf.CPHF(solver)
    op_objs = []
    for o in ops:
        op = operators.Operator(
            label=o.get("label", "dipole"),
            is_imaginary=o["is_imaginary"],
            is_spin_dependent=False,
            triplet=o.get("triplet", False),
            ao_integrals=o["ao_integrals"],
        )
        driver.add_operator(op)
        op_objs.append(op)
    driver.set_frequencies([float(f) for f in freqs])
    driver.run(hamiltonian=HAMS[ham], spin=SPINS[spin], program=None, program_obj=None)
    return solver, driver, op_objs


# ------------------------------------------------------------------ water / STO-3G (config 1)
def golden_water():
    C = utils.fix_mocoeffs_shape(utils.np_load_2(REFDIR / "C.npz"))
    E = utils.fix_moenergies_shape(utils.np_load_2(REFDIR / "F_MO.npz"))
    TEI = utils.np_load(REFDIR / "TEI_MO.npz")
    occ = (5, 2, 5, 2)
    mu = np.stack(
        [utils.parse_int_file_2(REFDIR / f"h2o_sto3g_mu{x}.dat", 7) for x in "xyz"], axis=0
    )
    freqs = (0.0, 0.02, 0.06, 0.1)
    out = {"C": C, "E": E, "TEI_MO": TEI, "dipole": mu, "frequencies": np.array(freqs)}
    for ham in HAMS:
        for spin in SPINS:
            for chol in (False, True):
                solver, driver, ops = run_reference(
                    C, E, occ, (TEI,), AO2MOTransformationType.full,
                    [{"is_imaginary": False, "ao_integrals": mu}], freqs, ham, spin, chol,
                )
                tag = f"{ham}_{spin}_{'chol' if chol else 'inv'}"
                out[f"results_{tag}"] = np.stack(driver.results)
                out[f"uncoupled_{tag}"] = np.stack(driver.uncoupled_results)
                out[f"rspvecs_{tag}"] = np.stack(ops[0].rspvecs_alph)
                if not chol:
                    out[f"hessian_last_{tag}"] = solver.explicit_hessian
    np.savez_compressed(HERE / "water_sto3g.npz", **out)
    print("water:", {k: v.shape for k, v in out.items() if k.startswith("results")})


# ------------------------------------------------------------------ LiH disk fixtures
def _load_lih(case, uhf):
    d = REFDIR / case
    occ = utils.read_file_occupations(d / "occupations")
    out = {
        "occupations": np.array(occ),
        "C": utils.read_file_3(d / "C"),
        "moene": utils.read_file_2(d / "moene"),
    }
    names = (
        ["moints_iajb_aaaa", "moints_iajb_aabb", "moints_iajb_bbaa", "moints_iajb_bbbb",
         "moints_ijab_aaaa", "moints_ijab_bbbb"]
        if uhf
        else ["moints_iajb_aaaa", "moints_ijab_aaaa"]
    )
    for nm in names:
        out[nm] = utils.read_file_4(d / nm)
    entries = []
    with open(d / "ref") as fh:
        for line in fh:
            t = line.split()
            assert len(t) == 6
            entries.append(t)
    labels = sorted({e[3] for e in entries} | {e[4] for e in entries})
    meta = {}
    for lab in labels:
        if "_spnorb" in lab:  # excluded by the reference too (tests/test_runners.py:34)
            continue
        op = dalton_label_to_operator(lab)
        meta[lab] = (op.label, op.slice_idx, op.is_imaginary, op.is_spin_dependent,
                     getattr(op, "hsofac", None))
        key = f"operator_mn_{op.label}"
        if key not in out:
            out[key] = utils.read_file_3(d / key)
    out["ref_hamiltonian"] = np.array([e[0] for e in entries])
    out["ref_spin"] = np.array([e[1] for e in entries])
    out["ref_frequency"] = np.array([e[2] for e in entries])
    out["ref_label_1"] = np.array([e[3] for e in entries])
    out["ref_label_2"] = np.array([e[4] for e in entries])
    out["ref_value"] = np.array([float(e[5]) for e in entries])
    out["label_names"] = np.array(list(meta.keys()))
    out["label_operator"] = np.array([meta[k][0] for k in meta])
    out["label_slice"] = np.array([meta[k][1] for k in meta])
    out["label_imaginary"] = np.array([meta[k][2] for k in meta])
    out["label_spin_dependent"] = np.array([meta[k][3] for k in meta])
    return out


def golden_lih():
    """Inputs + DALTON values + the reference's own answers for every non-excluded entry."""
    sys.modules.setdefault("cclib", types.ModuleType("cclib"))
    for case, uhf in (("r_lih_hf_sto-3g", False), ("u_lih_cation_hf_sto-3g", True)):
        out = _load_lih(case, uhf)
        occ = tuple(int(x) for x in out["occupations"])
        C = out["C"]
        if uhf:
            E = np.stack((np.diag(out["moene"][:, 0]), np.diag(out["moene"][:, 1])), axis=0)
            tei = tuple(out[k] for k in ("moints_iajb_aaaa", "moints_iajb_aabb",
                                         "moints_iajb_bbaa", "moints_iajb_bbbb",
                                         "moints_ijab_aaaa", "moints_ijab_bbbb"))
        else:
            E = np.diag(out["moene"][:, 0])[None]
            tei = (out["moints_iajb_aaaa"], out["moints_ijab_aaaa"])
        names = list(out["label_names"])
        vals = np.full(len(out["ref_value"]), np.nan)
        for k in range(len(vals)):
            l1, l2 = str(out["ref_label_1"][k]), str(out["ref_label_2"][k])
            if l1 not in names or l2 not in names:
                continue
            ops = []
            for lab in (l1, l2):
                q = names.index(lab)
                ints = out[f"operator_mn_{out['label_operator'][q]}"][int(out["label_slice"][q])]
                ops.append({"label": str(out["label_operator"][q]),
                            "is_imaginary": bool(out["label_imaginary"][q]),
                            "ao_integrals": ints[None]})
            _, driver, _ = run_reference(
                C, E, occ, tei, AO2MOTransformationType.partial, ops,
                [float(out["ref_frequency"][k])],
                str(out["ref_hamiltonian"][k]), str(out["ref_spin"][k]),
            )
            vals[k] = driver.results[0][1, 0]
        out["reference_value"] = vals
        np.savez_compressed(HERE / f"{case}.npz", **out)
        ok = ~np.isnan(vals)
        dev = np.abs(np.abs(out["ref_value"][ok]) - np.abs(vals[ok]))
        print(case, "entries", ok.sum(), "max |dalton|-|reference|", dev.max())


# ------------------------------------------------------------------ seeded synthetic cases
class _DenseJK(JK):
    """Dense stand-in for the psi4 JK object (contract: integrals.py:41-67)."""

    def __init__(self, I):
        super().__init__()
        self.I = I

    def compute_from_density(self, D):
        return np.einsum("pqrs,rs->pq", self.I, D), np.einsum("prqs,rs->pq", self.I, D)

    def compute_from_mocoeffs(self, C_left, C_right=None):
        if C_right is None:
            C_right = C_left
        return self.compute_from_density(np.dot(C_left, C_right.T))


def golden_synthetic_rhf(n=10, nocc=3):
    case = synthetic.rhf_case(n, nocc)
    C, E, occ, I = case["C"], case["E"], case["occupations"], case["I"]
    a = AO2MO(C, occ, I=I)
    a.perform_rhf_partial()
    ovov, oovv = a.tei_mo
    a.perform_rhf_full()
    (full,) = a.tei_mo
    Or = synthetic.operator_integrals(n, 3, imaginary=False)
    Oi = synthetic.operator_integrals(n, 3, imaginary=True)
    freqs = (0.0, 0.05, 0.1)
    out = {"n": n, "nocc": nocc, "ovov": ovov, "oovv": oovv, "full": full,
           "frequencies": np.array(freqs)}
    E0 = E[0]
    out["A_singlet_partial"] = eqp.form_rpa_a_matrix_mo_singlet_partial(E0, ovov, oovv)
    out["A_triplet_partial"] = eqp.form_rpa_a_matrix_mo_triplet_partial(E0, oovv)
    out["B_singlet_partial"] = eqp.form_rpa_b_matrix_mo_singlet_partial(ovov)
    out["B_triplet_partial"] = eqp.form_rpa_b_matrix_mo_triplet_partial(ovov)
    out["A_singlet_ss_partial"] = eqp.form_rpa_a_matrix_mo_singlet_ss_partial(E0, ovov, oovv)
    out["A_singlet_os_partial"] = eqp.form_rpa_a_matrix_mo_singlet_os_partial(ovov)
    out["B_singlet_ss_partial"] = eqp.form_rpa_b_matrix_mo_singlet_ss_partial(ovov)
    out["B_singlet_os_partial"] = eqp.form_rpa_b_matrix_mo_singlet_os_partial(ovov)
    out["A_singlet_full"] = eqf.form_rpa_a_matrix_mo_singlet_full(E0, full, nocc)
    out["A_triplet_full"] = eqf.form_rpa_a_matrix_mo_triplet_full(E0, full, nocc)
    out["B_singlet_full"] = eqf.form_rpa_b_matrix_mo_singlet_full(full, nocc)
    out["B_triplet_full"] = eqf.form_rpa_b_matrix_mo_triplet_full(full, nocc)
    out["A_singlet_ss_full"] = eqf.form_rpa_a_matrix_mo_singlet_ss_full(E0, full, nocc)
    out["A_singlet_os_full"] = eqf.form_rpa_a_matrix_mo_singlet_os_full(full, nocc, nocc)
    out["B_singlet_ss_full"] = eqf.form_rpa_b_matrix_mo_singlet_ss_full(full, nocc)
    out["B_singlet_os_full"] = eqf.form_rpa_b_matrix_mo_singlet_os_full(full, nocc, nocc)
    ops = [{"label": "dipole", "is_imaginary": False, "ao_integrals": Or},
           {"label": "angmom", "is_imaginary": True, "ao_integrals": Oi}]
    for ham in HAMS:
        for spin in SPINS:
            for ttype, tei in (("partial", (ovov, oovv)), ("full", (full,))):
                solver, driver, op_objs = run_reference(
                    C, E, occ, tei, AO2MOTransformationType[ttype], ops, freqs, ham, spin)
                tag = f"{ham}_{spin}_{ttype}"
                out[f"results_{tag}"] = np.stack(driver.results)
                out[f"uncoupled_{tag}"] = np.stack(driver.uncoupled_results)
                out[f"rspvecs0_{tag}"] = np.stack(op_objs[0].rspvecs_alph)
                out[f"rspvecs1_{tag}"] = np.stack(op_objs[1].rspvecs_alph)
                if ttype == "partial":
                    out[f"hessian_last_{tag}"] = solver.explicit_hessian
                    out[f"hessian_inv_last_{tag}"] = solver.explicit_hessian_inv
    out["rhs0_ai"] = op_objs[0].mo_integrals_ai_alph
    out["rhs0_super"] = op_objs[0].mo_integrals_ai_supervector_alph
    out["rhs1_super"] = op_objs[1].mo_integrals_ai_supervector_alph
    # Cholesky solver (untested in the reference)
    solver, driver, op_objs = run_reference(
        C, E, occ, (ovov, oovv), AO2MOTransformationType.partial, ops, freqs, "rpa", "singlet",
        cholesky=True)
    out["results_rpa_singlet_partial_chol"] = np.stack(driver.results)
    # iterative solver with a dense JK stand-in (real operator only, as tests/test_solvers.py)
    it = solvers.IterativeLinEqSolver(C, E, occ, _DenseJK(I), maxiter=200, conv=1e-24)
    drv = cphf.CPHF(it)
    op = operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or)
    drv.add_operator(op)
    drv.set_frequencies([0.0, 0.05])
    drv.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    out["iterative_results"] = np.stack(drv.results)
    out["iterative_rspvecs"] = np.stack(op.rspvecs_alph)
    out["iterative_frequencies"] = np.array([0.0, 0.05])
    np.savez_compressed(HERE / "synthetic_rhf_n10.npz", **out)
    print("synthetic rhf: static alpha", np.diag(out["results_rpa_singlet_partial"][0])[:3])


# ------------------------------------------------------------------ excitation energies
def _run_td(C, E, occ, tei, ttype, ops, ham, spin, tda_solver):
    from pymolresponse import td
    from pymolresponse.properties.ecd import ECD

    cls = solvers.ExactDiagonalizationSolverTDA if tda_solver else solvers.ExactDiagonalizationSolver
    solver = cls(C, E, occ)
    solver.tei_mo = tei
    solver.tei_mo_type = ttype
    driver = (td.TDA if tda_solver else td.TDHF)(solver)
    op_objs = []
    for o in ops:
        op = operators.Operator(label=o["label"], is_imaginary=o["is_imaginary"],
                                is_spin_dependent=False, triplet=False, ao_integrals=o["ao_integrals"])
        driver.add_operator(op)
        op_objs.append(op)
    driver.run(hamiltonian=HAMS[ham], spin=SPINS[spin], program=None, program_obj=None)
    ecd = ECD(None, None, driver, do_dipvel=False)
    ecd.form_results()
    return solver, op_objs, ecd


def golden_td():
    """Eigenvalues, oscillator strengths and rotational strengths from the reference's
    ExactDiagonalizationSolver(TDA) + td.TDHF / td.TDA + properties.ecd.ECD.form_results on the
    water/STO-3G fixture (full MO integrals) and the seeded synthetic n=10 case (partial)."""
    import contextlib
    import io

    out = {}
    Cw = utils.fix_mocoeffs_shape(utils.np_load_2(REFDIR / "C.npz"))
    Ew = utils.fix_moenergies_shape(utils.np_load_2(REFDIR / "F_MO.npz"))
    TEI = utils.np_load(REFDIR / "TEI_MO.npz")
    mu = np.stack([utils.parse_int_file_2(REFDIR / f"h2o_sto3g_mu{x}.dat", 7) for x in "xyz"], axis=0)
    Lw = synthetic.operator_integrals(7, 3, imaginary=True, seed=91)   # stands for angmom
    case = synthetic.rhf_case(10, 3)
    a = AO2MO(case["C"], case["occupations"], I=case["I"])
    a.perform_rhf_partial()
    Or = synthetic.operator_integrals(10, 3, imaginary=False)
    Oi = synthetic.operator_integrals(10, 3, imaginary=True)
    systems = {
        "water": (Cw, Ew, (5, 2, 5, 2), (TEI,), AO2MOTransformationType.full, Lw, mu),
        "syn10": (case["C"], case["E"], case["occupations"], a.tei_mo,
                  AO2MOTransformationType.partial, Oi, Or),
    }
    out["water_angmom_like"] = Lw
    # LiH / STO-3G disk fixture (real angular-momentum and dipole integrals, degenerate pi roots)
    lih = np.load(HERE / "r_lih_hf_sto-3g.npz")
    systems["lih"] = (lih["C"], np.diag(lih["moene"][:, 0])[None],
                      tuple(int(x) for x in lih["occupations"]),
                      (lih["moints_iajb_aaaa"], lih["moints_ijab_aaaa"]),
                      AO2MOTransformationType.partial, lih["operator_mn_angmom"],
                      lih["operator_mn_dipole"])
    for name, (C, E, occ, tei, ttype, Oimag, Oreal) in systems.items():
        ops = [{"label": "angmom", "is_imaginary": True, "ao_integrals": Oimag},
               {"label": "dipole", "is_imaginary": False, "ao_integrals": Oreal}]
        for tda_solver, ham in ((False, "rpa"), (False, "tda"), (True, "tda")):
            for spin in SPINS:
                with contextlib.redirect_stdout(io.StringIO()):
                    solver, op_objs, ecd = _run_td(C, E, occ, tei, ttype, ops, ham, spin, tda_solver)
                tag = f"{name}_{'tdasolver' if tda_solver else 'rpasolver'}_{ham}_{spin}"
                assert np.abs(solver.eigvals.imag).max() == 0.0
                out[f"eigvals_{tag}"] = solver.eigvals.real
                out[f"eigvecs_{tag}"] = solver.eigvecs.real
                for op in op_objs:
                    out[f"tmom_{op.label}_{tag}"] = op.transition_moments
                    out[f"osc_{op.label}_{tag}"] = op.oscillator_strengths
                    out[f"totosc_{op.label}_{tag}"] = op.total_oscillator_strengths
                out[f"rotstr_{tag}"] = ecd.rotational_strengths_diplen
    np.savez_compressed(HERE / "td_golden.npz", **out)
    print("td: water RPA singlet roots", out["eigvals_water_rpasolver_rpa_singlet"][:4])


def golden_synthetic_uhf(n=9, na=3, nb=2):
    case = synthetic.uhf_case(n, na, nb)
    C, E, occ, I = case["C"], case["E"], case["occupations"], case["I"]
    Coa, Cva, Cob, Cvb = C[0][:, :na], C[0][:, na:], C[1][:, :nb], C[1][:, nb:]
    T = AO2MO.transform
    # tuple orders of interfaces/pyscf/ao2mo.py:69 and :102-109, produced with the reference's
    # own NumPy transform (tests/test_ao2mo.py pins pyscf.ao2mo to it at 1e-15)
    partial = (T(I, Coa, Cva, Coa, Cva), T(I, Coa, Cva, Cob, Cvb), T(I, Cob, Cvb, Coa, Cva),
               T(I, Cob, Cvb, Cob, Cvb), T(I, Coa, Coa, Cva, Cva), T(I, Cob, Cob, Cvb, Cvb))
    full = (T(I, C[0], C[0], C[0], C[0]), T(I, C[0], C[0], C[1], C[1]),
            T(I, C[1], C[1], C[0], C[0]), T(I, C[1], C[1], C[1], C[1]))
    Or = synthetic.operator_integrals(n, 3, imaginary=False)
    Oi = synthetic.operator_integrals(n, 3, imaginary=True)
    ops = [{"label": "dipole", "is_imaginary": False, "ao_integrals": Or},
           {"label": "angmom", "is_imaginary": True, "ao_integrals": Oi}]
    freqs = (0.0, 0.07)
    out = {"n": n, "na": na, "nb": nb, "frequencies": np.array(freqs)}
    for k, t in enumerate(partial):
        out[f"partial_{k}"] = t
    for k, t in enumerate(full):
        out[f"full_{k}"] = t
    out["A_os_a_full"] = eqf.form_rpa_a_matrix_mo_singlet_os_full(full[1], na, nb)
    out["B_os_b_full"] = eqf.form_rpa_b_matrix_mo_singlet_os_full(full[2], nb, na)
    for ham in HAMS:
        for spin in SPINS:
            for ttype, tei in (("partial", partial), ("full", full)):
                solver, driver, op_objs = run_reference(
                    C, E, occ, tei, AO2MOTransformationType[ttype], ops, freqs, ham, spin)
                tag = f"{ham}_{spin}_{ttype}"
                out[f"results_{tag}"] = np.stack(driver.results)
                out[f"uncoupled_{tag}"] = np.stack(driver.uncoupled_results)
                out[f"rspvecs0_alph_{tag}"] = np.stack(op_objs[0].rspvecs_alph)
                out[f"rspvecs0_beta_{tag}"] = np.stack(op_objs[0].rspvecs_beta)
                out[f"rspvecs1_alph_{tag}"] = np.stack(op_objs[1].rspvecs_alph)
                out[f"rspvecs1_beta_{tag}"] = np.stack(op_objs[1].rspvecs_beta)
                if ttype == "partial" and ham == "rpa" and spin == "singlet":
                    for q, g in enumerate(solver.explicit_hessian):
                        out[f"hessian_last_{q}"] = g
                    out["left_alph_last"] = solver.left_alph
                    out["left_beta_last"] = solver.left_beta
    out["rhs0_super_alph"] = op_objs[0].mo_integrals_ai_supervector_alph
    out["rhs0_super_beta"] = op_objs[0].mo_integrals_ai_supervector_beta
    solver, driver, op_objs = run_reference(
        C, E, occ, partial, AO2MOTransformationType.partial, ops, freqs, "rpa", "singlet",
        cholesky=True)
    out["results_rpa_singlet_partial_chol"] = np.stack(driver.results)
    np.savez_compressed(HERE / "synthetic_uhf_n9.npz", **out)
    print("synthetic uhf: static", np.diag(out["results_rpa_singlet_partial"][0])[:3])


if __name__ == "__main__":
    golden_water()
    golden_synthetic_rhf()
    golden_synthetic_uhf()
    golden_lih()
    golden_td()
