"""Base classes of the property wrappers (reference ``pymolresponse/molecular_property.py``)."""

from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Sequence
from typing import Any

from pymolresponse_b200.cphf import CPHF
from pymolresponse_b200.core import Program
from pymolresponse_b200.integrals import Integrals


class MolecularProperty(ABC):
    """A molecular property calculated from one or more operators."""

    def __init__(self, program, program_obj: Any, driver) -> None:
        self.program = program
        self.program_obj = program_obj
        self.driver = driver

    def run(self, hamiltonian, spin) -> None:
        assert self.driver.solver is not None
        self.driver.run(hamiltonian=hamiltonian, spin=spin, program=self.program,
                        program_obj=self.program_obj)

    @abstractmethod
    def form_operators(self) -> None:
        pass

    @abstractmethod
    def form_results(self) -> None:
        pass

    def _integral_generator(self):
        """Pick the AO-integral provider.  An :class:`Integrals` instance passed as ``program_obj``
        (``Program.Arrays``) needs no quantum-chemistry package; pyscf / psi4 objects go through
        thin adapters that are imported lazily (reference properties/electric.py:29-38)."""
        if isinstance(self.program_obj, Integrals):
            return self.program_obj
        if self.program == Program.PySCF:
            from pymolresponse_b200.interfaces.pyscf_adapter import IntegralsPyscf

            return IntegralsPyscf(self.program_obj)
        if self.program == Program.Psi4:
            from pymolresponse_b200.interfaces.psi4_adapter import IntegralsPsi4

            return IntegralsPsi4(self.program_obj)
        raise RuntimeError


class ResponseProperty(MolecularProperty, ABC):
    """A property that is the response to one or more perturbations."""

    def __init__(self, program, program_obj: Any, driver: CPHF, *,
                 frequencies: Sequence[float]) -> None:
        driver.set_frequencies(frequencies)
        super().__init__(program, program_obj, driver)
        assert isinstance(self.driver, CPHF)
        self.frequencies = self.driver.frequencies


class TransitionProperty(MolecularProperty, ABC):
    """A property built from excited-state transition moments (reference
    molecular_property.py:62-65)."""

    def __init__(self, program, program_obj: Any, driver) -> None:
        from pymolresponse_b200.td import TDHF

        super().__init__(program, program_obj, driver)
        assert isinstance(self.driver, TDHF)
