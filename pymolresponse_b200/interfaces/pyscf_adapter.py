"""pyscf adapters (lazy import; pyscf only *generates* AO integrals, it is not on the hot path).
Reference: ``pymolresponse/interfaces/pyscf/integrals.py:13-36``, ``utils.py:13-17``."""

from __future__ import annotations

from pymolresponse_b200 import integrals


class IntegralsPyscf(integrals.Integrals):
    _INTOR = {
        "dipole": ("cint1e_r_sph", 3),
        "angmom_common_gauge": ("cint1e_cg_irxp_sph", 3),
        "angmom_giao": ("cint1e_giao_irjxp_sph", 3),
    }

    def __init__(self, pyscfmol) -> None:
        super().__init__()
        self.mol = pyscfmol

    def _compute(self, label):
        name, comp = self._INTOR[label.label]
        return self.mol.intor(name, comp=label.comp or comp)


def occupations_from_pyscf_mol(mol, C):
    norb = C.shape[-1]
    nocc_a, nocc_b = mol.nelec
    return (nocc_a, norb - nocc_a, nocc_b, norb - nocc_b)
