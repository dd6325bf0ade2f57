"""Optical rotatory dispersion (reference ``pymolresponse/properties/optrot.py:13-96``)."""

from __future__ import annotations

from collections.abc import Sequence
from typing import Any

from pymolresponse_b200 import integrals
from pymolresponse_b200.molecular_property import ResponseProperty
from pymolresponse_b200.operators import Operator


class ORD(ResponseProperty):
    """Mixed electric-dipole / magnetic-dipole polarizability G'(w): the imaginary angular momentum
    and the real dipole operator (and optionally the imaginary velocity-gauge dipole) solved as ONE
    batch of 6 (9) right-hand sides per frequency; operator order angmom, dipole, dipvel."""

    def __init__(self, program, program_obj: Any, driver, *,
                 frequencies: Sequence[float] = [0.0], do_dipvel: bool = False) -> None:  # noqa: B006
        super().__init__(program, program_obj, driver, frequencies=frequencies)
        self.do_dipvel = do_dipvel

    def form_operators(self) -> None:
        generator = self._integral_generator()
        self.driver.add_operator(Operator(
            label="angmom", is_imaginary=True, is_spin_dependent=False, triplet=False,
            ao_integrals=generator.integrals(integrals.ANGMOM_COMMON_GAUGE)))
        self.driver.add_operator(Operator(
            label="dipole", is_imaginary=False, is_spin_dependent=False, triplet=False,
            ao_integrals=generator.integrals(integrals.DIPOLE)))
        if self.do_dipvel:
            self.driver.add_operator(Operator(
                label="dipvel", is_imaginary=True, is_spin_dependent=False, triplet=False,
                ao_integrals=generator.integrals(integrals.DIPVEL)))

    def form_results(self) -> None:
        assert len(self.driver.results) == len(self.frequencies)
        self.polarizabilities = []
        self.polarizabilities_lenmag = []
        for idxf, _frequency in enumerate(self.frequencies):
            results = self.driver.results[idxf]
            assert results.shape == ((9, 9) if self.do_dipvel else (6, 6))
            self.polarizabilities.append(results[3:6, 3:6])
            self.polarizabilities_lenmag.append(results[0:3, 3:6])
