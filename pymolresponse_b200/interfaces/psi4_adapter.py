"""psi4 adapter (lazy import).  Reference: ``pymolresponse/interfaces/psi4/integrals.py:20-140``."""

from __future__ import annotations

import numpy as np

from pymolresponse_b200 import integrals


class IntegralsPsi4(integrals.Integrals):
    def __init__(self, molecule) -> None:
        import psi4

        super().__init__()
        wfn = psi4.core.Wavefunction.build(molecule, psi4.core.get_global_option("BASIS"))
        self._mints = psi4.core.MintsHelper(wfn.basisset())

    def _compute(self, label):
        if label.label == "dipole":
            return np.stack([np.asarray(m) for m in self._mints.ao_dipole()])
        if label.label.startswith("angmom"):
            return np.stack([np.asarray(m) for m in self._mints.ao_angular_momentum()])
        raise KeyError(label)
