"""Inputs of the excitation-energy golden cases (tests/golden/td_golden.npz, written by
tests/golden/make_golden.py golden_td from the unmodified reference)."""

import numpy as np

TD_COMBOS = [(False, "rpa"), (False, "tda"), (True, "tda")]   # (tda_solver, hamiltonian)
ESUECD_TOL = 1e-6


def td_system(golden_dir, name):
    """-> dict(C, E, occ, tei, ttype, Oimag, Oreal)."""
    from pymolresponse_b200 import synthetic

    td = np.load(golden_dir / "td_golden.npz")
    if name == "water":
        g = np.load(golden_dir / "water_sto3g.npz")
        return {"C": g["C"], "E": g["E"], "occ": (5, 2, 5, 2), "tei": (g["TEI_MO"],), "ttype": "full",
                "Oimag": td["water_angmom_like"], "Oreal": g["dipole"]}
    if name == "lih":
        g = np.load(golden_dir / "r_lih_hf_sto-3g.npz")
        return {"C": g["C"], "E": np.diag(g["moene"][:, 0])[None],
                "occ": tuple(int(x) for x in g["occupations"]),
                "tei": (g["moints_iajb_aaaa"], g["moints_ijab_aaaa"]), "ttype": "partial",
                "Oimag": g["operator_mn_angmom"], "Oreal": g["operator_mn_dipole"]}
    g = np.load(golden_dir / "synthetic_rhf_n10.npz")
    n, o = int(g["n"]), int(g["nocc"])
    case = synthetic.rhf_case(n, o, build_eri=False)
    return {"C": case["C"], "E": case["E"], "occ": case["occupations"], "tei": (g["ovov"], g["oovv"]),
            "ttype": "partial", "Oimag": synthetic.operator_integrals(n, 3, True),
            "Oreal": synthetic.operator_integrals(n, 3, False)}


def tag_of(name, tda_solver, ham, spin):
    return f"{name}_{'tdasolver' if tda_solver else 'rpasolver'}_{ham}_{spin}"
