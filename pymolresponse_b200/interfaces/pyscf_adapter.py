"""pyscf adapters (lazy import; pyscf only *generates* AO integrals, it is not on the hot path).
Reference: ``pymolresponse/interfaces/pyscf/integrals.py:13-36``, ``utils.py:13-17``."""

from __future__ import annotations

from pymolresponse_b200 import integrals


class IntegralsPyscf(integrals.Integrals):
    _INTOR = {
        "dipole": ("cint1e_r_sph", 3),
        "angmom_common_gauge": ("cint1e_cg_irxp_sph", 3),
        "angmom_giao": ("cint1e_giao_irjxp_sph", 3),
        "dipvel": ("cint1e_ipovlp_sph", 3),
    }

    def __init__(self, pyscfmol) -> None:
        super().__init__()
        self.mol = pyscfmol

    def _compute(self, label):
        if label.label in ("spinorb", "spinorb_eff"):
            return self._spin_orbit(effective=label.label == "spinorb_eff")
        name, comp = self._INTOR[label.label]
        return self.mol.intor(name, comp=label.comp or comp)

    def _spin_orbit(self, effective: bool):
        """sum_A Z_A * <p x 1/|r - R_A| p> (reference properties/magnetic.py:127-154; the
        effective charges are all zero there, :145)."""
        total = 0
        for atm_id in range(self.mol.natm):
            self.mol.set_rinv_orig(self.mol.atom_coord(atm_id))
            chg = 0 if effective else self.mol.atom_charge(atm_id)
            total = total + chg * self.mol.intor("cint1e_prinvxp_sph", comp=3)
        return total


def occupations_from_pyscf_mol(mol, C):
    norb = C.shape[-1]
    nocc_a, nocc_b = mol.nelec
    return (nocc_a, norb - nocc_a, nocc_b, norb - nocc_b)


def calculate_origin_pyscf(origin_string: str, mol, mocoeffs, occupations):
    """Gauge origin keywords of the reference (interfaces/pyscf/helpers.py:99-150): zero, center of
    mass, of nuclear charge, of electronic charge."""
    import numpy as np

    key = origin_string.lower()
    coords = mol.atom_coords()
    charges = np.asarray(mol.atom_charges(), dtype=float)
    if key in ("explicitly-set", "zero"):
        return np.zeros(3)
    if key in ("com", "centerofmass"):
        masses = np.asarray(mol.atom_mass_list(isotope_avg=False), dtype=float)
        return (masses[:, None] * coords).sum(axis=0) / masses.sum()
    if key in ("ncc", "centerofnuccharge"):
        return (charges[:, None] * coords).sum(axis=0) / charges.sum()
    if key in ("ecc", "centerofelcharge"):
        nocc_a, _, nocc_b, _ = occupations
        if mocoeffs.shape[0] == 2:
            D = (mocoeffs[0][:, :nocc_a] @ mocoeffs[0][:, :nocc_a].T
                 + mocoeffs[1][:, :nocc_b] @ mocoeffs[1][:, :nocc_b].T)
        else:
            D = 2 * mocoeffs[0][:, :nocc_a] @ mocoeffs[0][:, :nocc_a].T
        mol.set_common_orig([0.0, 0.0, 0.0])
        dip = mol.intor("cint1e_r_sph", comp=3)
        return np.einsum("xpq,pq->x", dip, D) / np.trace(D @ mol.intor("int1e_ovlp"))
    raise ValueError(f"unknown gauge origin keyword: {origin_string}")
