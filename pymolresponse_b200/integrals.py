"""Integral-provider interfaces (reference ``pymolresponse/integrals.py:15-67``) plus the
array-backed and GPU-resident implementations the hot path uses.

``Integrals`` / ``JK`` keep the reference's abstract contracts.  New here:

* :class:`IntegralsArrays` -- property integrals supplied as NumPy arrays keyed by label, so the
  property wrappers run without pyscf/psi4;
* :class:`JKDense` -- J/K builds against a dense AO ERI tensor held in HBM, every requested density
  contracted in ONE pass of ``pmr_jk_contract`` (the reference's only working JK is psi4's,
  interfaces/psi4/integrals.py:143-181).
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Mapping
from dataclasses import dataclass
from enum import Enum, auto, unique

import numpy as np

from pymolresponse_b200 import _native as nv


@unique
class IntegralSymmetry(Enum):
    SYMMETRIC = auto()
    ANTISYMMETRIC = auto()


@dataclass(frozen=True)
class IntegralLabel:
    label: str
    comp: int | None = None
    symmetry: IntegralSymmetry = IntegralSymmetry.SYMMETRIC


# labels the property wrappers ask for (names follow interfaces/pyscf/integrals.py)
DIPOLE = IntegralLabel("dipole", 3, IntegralSymmetry.SYMMETRIC)
ANGMOM_COMMON_GAUGE = IntegralLabel("angmom_common_gauge", 3, IntegralSymmetry.ANTISYMMETRIC)
ANGMOM_GIAO = IntegralLabel("angmom_giao", 3, IntegralSymmetry.ANTISYMMETRIC)
DIPVEL = IntegralLabel("dipvel", 3, IntegralSymmetry.ANTISYMMETRIC)
# one-electron spin-orbit integrals summed over the nuclei (charge-weighted), see
# properties/magnetic.py ElectronicGTensor.form_operators
SO_SPHER_1e = IntegralLabel("spinorb", 3, IntegralSymmetry.ANTISYMMETRIC)
SO_SPHER_1e_EFF = IntegralLabel("spinorb_eff", 3, IntegralSymmetry.ANTISYMMETRIC)


class Integrals(ABC):
    def __init__(self) -> None:
        pass

    def integrals(self, label: IntegralLabel):
        return self._compute(label)

    @abstractmethod
    def _compute(self, label: IntegralLabel):
        """Compute the integrals of some operator using some implementation."""


class IntegralsArrays(Integrals):
    """AO property integrals the caller already holds: ``{label or label.label: (ncomp,n,n)}``."""

    def __init__(self, arrays: Mapping) -> None:
        super().__init__()
        self.arrays = dict(arrays)

    def _compute(self, label: IntegralLabel):
        for key in (label, getattr(label, "label", None)):
            if key in self.arrays:
                return self.arrays[key]
        raise KeyError(f"no integrals stored for {label}")


class JK(ABC):
    """Coulomb and exchange matrices pre-contracted with a density (reference integrals.py:41-67):
    ``J_pq = sum (pq|rs) D_rs``, ``K_pq = sum (pr|qs) D_rs`` with ``D = C_left . C_right^T``."""

    def __init__(self) -> None:
        pass

    @abstractmethod
    def compute_from_density(self, D):
        """Compute J and K from some density."""

    @abstractmethod
    def compute_from_mocoeffs(self, C_left, C_right=None):
        """Compute J and K from MO coefficients."""

    def compute_from_mocoeffs_batch(self, lefts, rights):
        """Several (C_left, C_right) pairs at once -> list of (J, K) CUDA tensors.  The default
        loops over :meth:`compute_from_mocoeffs` on the host interface (any user JK works)."""
        out = []
        for cl, cr in zip(lefts, rights):
            J, K = self.compute_from_mocoeffs(nv.to_host(cl), nv.to_host(cr))
            out.append((nv.to_dev(J), nv.to_dev(K)))
        return out


class JKDense(JK):
    """J/K against a dense (n,n,n,n) AO ERI tensor resident in HBM (``pmr_jk_contract``)."""

    MAX_DENSITIES = 8

    def __init__(self, I) -> None:  # noqa: E741
        super().__init__()
        self.I = nv.to_dev(I)
        assert self.I.dim() == 4

    def _contract(self, D):
        """D: (nd, n, n) CUDA tensor -> (J, K) of the same shape."""
        n = int(self.I.shape[0])
        nd = int(D.shape[0])
        J, K = nv.empty(nd, n, n), nv.empty(nd, n, n)
        L = nv.lib()
        for d0 in range(0, nd, self.MAX_DENSITIES):
            d1 = min(nd, d0 + self.MAX_DENSITIES)
            nv.check(L.pmr_jk_contract(nv.ptr(self.I), n, nv.ptr(D, d0 * n * n), d1 - d0,
                                       nv.ptr(J, d0 * n * n), nv.ptr(K, d0 * n * n),
                                       nv.stream_ptr()))
        return J, K

    def compute_from_density(self, D):
        on_dev = nv.is_device_array(D)
        J, K = self._contract(nv.to_dev(D)[None].contiguous())
        return (J[0], K[0]) if on_dev else (nv.to_host(J[0]), nv.to_host(K[0]))

    def compute_from_mocoeffs(self, C_left, C_right=None):
        if C_right is None:
            C_right = C_left
        on_dev = nv.is_device_array(C_left)
        (J, K), = self.compute_from_mocoeffs_batch([nv.to_dev(C_left)], [nv.to_dev(C_right)])
        return (J, K) if on_dev else (nv.to_host(J), nv.to_host(K))

    def compute_from_mocoeffs_batch(self, lefts, rights):
        n = int(self.I.shape[0])
        D = nv.empty(len(lefts), n, n)
        for k, (cl, cr) in enumerate(zip(lefts, rights)):
            D[k] = nv.matmul(nv.to_dev(cl), nv.to_dev(cr), transpose_b=True)
        J, K = self._contract(D)
        return [(J[k], K[k]) for k in range(len(lefts))]


_ = np
