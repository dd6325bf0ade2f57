"""Magnetic properties (reference ``pymolresponse/properties/magnetic.py:17-190``)."""

from __future__ import annotations

from typing import Any

import numpy as np

from pymolresponse_b200 import integrals
from pymolresponse_b200.molecular_property import ResponseProperty
from pymolresponse_b200.operators import Operator


class Magnetizability(ResponseProperty):
    """Paramagnetic magnetizability: static response to the (imaginary) angular-momentum
    operator, ``xi = results[0] / 4``."""

    def __init__(self, program, program_obj: Any, driver, use_giao: bool = False) -> None:
        super().__init__(program, program_obj, driver, frequencies=np.asarray([0.0]))
        self.use_giao = use_giao

    def form_operators(self) -> None:
        generator = self._integral_generator()
        label = integrals.ANGMOM_GIAO if self.use_giao else integrals.ANGMOM_COMMON_GAUGE
        operator_angmom = Operator(
            label="angmom", is_imaginary=True, is_spin_dependent=False, triplet=False,
            ao_integrals=generator.integrals(label),
        )
        self.driver.add_operator(operator_angmom)

    def form_results(self) -> None:
        assert len(self.driver.results) == 1
        self.magnetizability = (1 / 4) * self.driver.results[0]


class ElectronicGTensor(ResponseProperty):
    """Orbital-Zeeman / spin-orbit part of the electronic g-tensor (reference
    properties/magnetic.py:57-190): three imaginary operators (angular momentum, one-electron
    spin-orbit with exact and with effective nuclear charges) in one static 9-column batch.

    ``gauge_origin`` is a 3-vector, or -- pyscf only -- one of the origin keywords the reference
    resolves through pyscf.  With ``Program.Arrays`` the integral provider already holds the
    integrals at the wanted origin and ``nelec`` gives (n_alpha, n_beta)."""

    def __init__(self, program, program_obj: Any, driver, *, gauge_origin="ecc",
                 nelec: tuple[int, int] | None = None) -> None:
        super().__init__(program, program_obj, driver, frequencies=[0.0])
        self.nelec = nelec
        if isinstance(gauge_origin, str):
            self.gauge_origin = gauge_origin if self._is_pyscf_mol() else None
        else:
            origin = np.asarray(gauge_origin, dtype=float).reshape(-1)
            assert origin.shape == (3,)
            self.gauge_origin = origin

    def _is_pyscf_mol(self) -> bool:
        return hasattr(self.program_obj, "set_common_orig")

    def _resolve_origin(self):
        if isinstance(self.gauge_origin, str):
            from pymolresponse_b200.interfaces.pyscf_adapter import calculate_origin_pyscf

            return calculate_origin_pyscf(self.gauge_origin, self.program_obj,
                                          self.driver.solver.mocoeffs, self.driver.solver.occupations)
        return self.gauge_origin

    def form_operators(self) -> None:
        generator = self._integral_generator()
        if self._is_pyscf_mol():
            self.program_obj.set_common_orig(self._resolve_origin())
        for label, ilabel in (("angmom", integrals.ANGMOM_COMMON_GAUGE),
                              ("spinorb", integrals.SO_SPHER_1e),
                              ("spinorb_eff", integrals.SO_SPHER_1e_EFF)):
            self.driver.add_operator(Operator(
                label=label, is_imaginary=True, is_spin_dependent=False, triplet=False,
                ao_integrals=generator.integrals(ilabel)))

    def form_results(self) -> None:
        assert len(self.driver.results) == 1
        results = self.driver.results[0]
        assert results.shape == (9, 9)
        block_2 = results[0:3, 3:6]  # angmom/spinorb
        nelec = self.nelec if self.nelec is not None else self.program_obj.nelec
        nalph, nbeta = nelec
        exact_spin = 0.5 * (nalph - nbeta)
        res_1 = block_2 / exact_spin
        # principal values are sqrt(eigvals(g.T g))
        self.g_oz_soc_1 = res_1
        self.g_oz_soc_1_eig = np.sqrt(np.linalg.eigvals(np.dot(res_1.T, res_1)))
