"""Magnetic properties (reference ``pymolresponse/properties/magnetic.py:17-54``)."""

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
