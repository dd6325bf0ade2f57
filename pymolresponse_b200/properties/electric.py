"""Electric properties (reference ``pymolresponse/properties/electric.py``)."""

from __future__ import annotations

from collections.abc import Sequence
from typing import Any

from pymolresponse_b200 import integrals
from pymolresponse_b200.molecular_property import ResponseProperty
from pymolresponse_b200.operators import Operator


class Polarizability(ResponseProperty):
    """Dipole polarizability alpha(w): one real 3-component operator, one (3,3) tensor per
    frequency (reference properties/electric.py:15-56)."""

    def __init__(self, program, program_obj: Any, driver, *,
                 frequencies: Sequence[float] = [0.0]) -> None:  # noqa: B006
        super().__init__(program, program_obj, driver, frequencies=frequencies)
        self.polarizabilities = []

    def form_operators(self) -> None:
        generator = self._integral_generator()
        operator_diplen = Operator(
            label="dipole", is_imaginary=False, is_spin_dependent=False, triplet=False,
            ao_integrals=generator.integrals(integrals.DIPOLE),
        )
        self.driver.add_operator(operator_diplen)

    def form_results(self) -> None:
        assert len(self.driver.results) == len(self.frequencies)
        for idxf, _frequency in enumerate(self.frequencies):
            results = self.driver.results[idxf]
            assert results.shape == (3, 3)
            self.polarizabilities.append(results)
