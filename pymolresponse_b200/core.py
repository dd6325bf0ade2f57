"""Enumerations shared by the drivers, solvers and integral transforms.

Same member names and values as the reference's ``pymolresponse/core.py:6-60`` so that
``Hamiltonian.RPA``, ``Spin.singlet``, ``AO2MOTransformationType.partial`` ... can be used
interchangeably (``Hamiltonian["RPA"]`` and ``Spin["singlet"]`` look-ups are part of the
reference's test harness, ``tests/test_calculators.py:76-77``).
"""

from enum import Enum, unique


@unique
class AO2MOTransformationType(Enum):
    """partial: (ia|jb) and (ij|ab) blocks only; full: the whole (pq|rs) tensor."""

    partial = "partial"
    full = "full"


@unique
class Hamiltonian(Enum):
    """RPA keeps the B block of the orbital Hessian; TDA sets it to zero."""

    RPA = "rpa"
    TDA = "tda"


@unique
class Spin(Enum):
    """Spin-conserving (singlet) or spin-flipping (triplet) response."""

    singlet = 0
    triplet = 1


@unique
class Program(Enum):
    """Which back end supplies integrals.  ``Arrays`` (new here) means "the caller already
    holds NumPy/CUDA arrays": no quantum-chemistry program is needed for the hot path."""

    PySCF = "pyscf"
    Psi4 = "psi4"
    DALTON = "dalton"
    Arrays = "arrays"
