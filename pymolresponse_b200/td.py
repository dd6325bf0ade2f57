"""Drivers for the time-dependent Hartree-Fock (TDHF / RPA) and Tamm-Dancoff (TDA / CIS)
excitation-energy equations.  Mirrors the reference's ``pymolresponse/td.py`` (``TDHF``, ``TDA``):
same attributes on the operators (``transition_moments``, ``oscillator_strengths``,
``total_oscillator_strengths``) and on the solver (``eigvecs_normed``).

The eigenvectors stay in HBM after the solver ran: normalisation is a column scaling on the device
and the transition moments of every operator with every state are ONE DMMA GEMM
``2 * P (ncomp x dim) . Z (dim x nstates)``; only the small (nstates, ncomp) tables come back.
"""

from __future__ import annotations

from typing import Any, ClassVar

import numpy as np

from pymolresponse_b200 import _native as nv
from pymolresponse_b200.constants import HARTREE_TO_EV, HARTREE_TO_INVCM
from pymolresponse_b200.core import Hamiltonian, Spin
from pymolresponse_b200.cphf import CPHF
from pymolresponse_b200.solvers import EigSolver, EigSolverTDA
from pymolresponse_b200.utils import form_indices_orbwin, form_vec_energy_differences


class TDHF(CPHF):
    """Driver for the TDHF (RPA) equations (reference td.py:20-77)."""

    #: which operator gradient is contracted with the eigenvectors ("super": [g; s g] rows)
    _gradient = "super"

    def __init__(self, solver: EigSolver) -> None:
        self.solver = solver

    def run(self, hamiltonian, spin, program, program_obj: Any) -> None:
        self.hamiltonian = hamiltonian
        self.spin = spin
        self.solver.run(hamiltonian, spin, program, program_obj)
        self.form_results()

    # ------------------------------------------------------------------ normalisation
    def _normed_eigvecs_dev(self):
        """Columns scaled so that 2 (|x|^2 - |y|^2) = 1 (EigSolver.norm_xy, solvers.py:685-690)."""
        Z = self.solver._eigvecs_dev
        dim, nstates = int(Z.shape[0]), int(Z.shape[1])
        N = dim // 2
        torch = nv.torch_mod()
        from pymolresponse_b200.utils import contract_rows

        Zt = Z.t().contiguous()                                     # rows = states
        ones = nv.to_dev(np.ones((1, N)))
        x2 = contract_rows(ones, (Zt[:, :N] * Zt[:, :N]).contiguous())[0]
        y2 = contract_rows(ones, (Zt[:, N:] * Zt[:, N:]).contiguous())[0]
        scale = (1.0 / torch.sqrt(2.0 * (x2 - y2))).contiguous()
        Zn = Z.clone()
        nv.check(nv.lib().pmr_scale_columns(nv.ptr(Zn), dim, nstates, nstates, nv.ptr(scale),
                                            nv.stream_ptr()))
        return Zn

    def _gradient_rows(self, operator):
        key = "super_alph" if self._gradient == "super" else "ai_alph"
        if key in operator._dev:
            return operator._dev[key]
        name = ("mo_integrals_ai_supervector_alph" if self._gradient == "super"
                else "mo_integrals_ai_alph")
        return nv.to_dev(getattr(operator, name))[:, :, 0].contiguous()

    def form_results(self) -> None:
        assert isinstance(self.solver, EigSolver)
        Zn = self._normed_eigvecs_dev()
        eigvals = np.asarray(nv.to_host(self.solver._eigvals_dev)).real
        for operator in self.solver.operators:
            P = self._gradient_rows(operator)
            tm = nv.to_host(nv.matmul(P, Zn, alpha=2.0)).T            # (nstates, ncomp)
            operator.transition_moments = tm
            operator.oscillator_strengths = (2 / 3) * eigvals[:, None] * tm**2
            operator.total_oscillator_strengths = (2 / 3) * eigvals * np.einsum("sc,sc->s", tm, tm)
        self.solver.eigvecs_normed = Zn if self.solver.device_eig_results else nv.to_host(Zn)

    # ------------------------------------------------------------------ reports
    def print_results(self) -> None:
        energies = np.asarray(self.solver.eigvals).real
        for idx in range(len(energies)):
            print("=" * 78)
            print(f" State: {idx + 1}")
            print(f" Excitation energy [a.u.]: {energies[idx]}")
            print(f" Excitation energy [eV]  : {energies[idx] * HARTREE_TO_EV}")
            for operator in self.solver.operators:
                print("-" * 78)
                print(f" Operator: {operator.label}")
                print(f" Transition moment: {operator.transition_moments[idx]}")
                print(f" Oscillator strength: {operator.oscillator_strengths[idx]}")
                print(f" Oscillator strength (total): {operator.total_oscillator_strengths[idx]}")

    def print_results_nwchem(self) -> str:
        """The '10 smallest eigenvalue differences' table (reference td.py:97-121)."""
        nocc_tot, nvirt_tot, _, _ = self.solver.occupations
        moene = np.diag(self.solver.moenergies[0])
        ediff = form_vec_energy_differences(moene[:nocc_tot], moene[nocc_tot:]) * HARTREE_TO_EV
        idxsort = np.argsort(ediff)
        ediff_sorted = ediff[idxsort]
        pairs = form_indices_orbwin(nocc_tot, nvirt_tot)
        indices_sorted = [pairs[i] for i in idxsort]
        ndiff = 10
        lines = [f"   {ndiff:>2d} smallest eigenvalue differences (eV) ",
                 "--------------------------------------------------------",
                 "  No. Spin  Occ  Vir  Irrep   E(Occ)    E(Vir)   E(Diff)",
                 "--------------------------------------------------------"]
        for idx in range(min(ndiff, len(indices_sorted))):
            iocc, ivirt = indices_sorted[idx]
            lines.append(
                f"{idx + 1:>5d}{1:>5d}{iocc + 1:>5d}{ivirt + 1:>5d} "
                f"{'X':<5}{moene[iocc]:>10.3f}{moene[ivirt]:>10.3f}{ediff_sorted[idx]:>10.3f}")
        lines.append("--------------------------------------------------------")
        return "\n".join(lines)

    _HAMILTONIAN_MAP_ORCA: ClassVar = {Hamiltonian.TDA: "CIS-", Hamiltonian.RPA: "RPA "}
    _SPIN_MAP_ORCA: ClassVar = {Spin.singlet: "SINGLETS", Spin.triplet: "TRIPLETS"}

    def print_results_orca(self, cutoff: float = 0.01) -> str:
        """Excited-state listing with the leading excitations (reference td.py:127-172)."""
        energies = np.asarray(self.solver.eigvals).real
        nocc_tot, nvirt_tot, _, _ = self.solver.occupations
        indices = form_indices_orbwin(nocc_tot, nvirt_tot)
        eigvecs = np.asarray(nv.to_host(self.solver.eigvecs))
        lines = [
            "-----------------------------",
            f"{self._HAMILTONIAN_MAP_ORCA[self.hamiltonian]}EXCITED STATES "
            f"({self._SPIN_MAP_ORCA[self.spin]})",
            "-----------------------------",
            "",
            f"the weight of the individual excitations are printed if larger than {cutoff:>4.2f}",
            "",
        ]
        for state in range(len(energies)):
            lines.append(
                f"STATE{state + 1:>3d}:  E={energies[state]:>11.6f} "
                f"au{energies[state] * HARTREE_TO_EV:>11.3f} eV"
                f"{energies[state] * HARTREE_TO_INVCM:>11.1f} cm**-1")
            vec = eigvecs[:, state]
            for p in np.nonzero(vec**2 > cutoff)[0]:
                iocc, ivirt = indices[p % len(indices)]
                lines.append(f"{iocc:>6d}a ->{ivirt:>4d}a  :{vec[p] ** 2:>13.6f} (c={vec[p]:>12.8f})")
            lines.append("")
        return "\n".join(lines)


class TDA(TDHF):
    """Driver for the TDA / CIS equations (reference td.py:190-241): eigenvectors scaled by
    1/sqrt(2), contracted with the plain ai gradients."""

    _gradient = "ai"

    def __init__(self, solver: EigSolverTDA) -> None:
        self.solver = solver

    def _normed_eigvecs_dev(self):
        Z = self.solver._eigvecs_dev
        from pymolresponse_b200 import _engine as eng

        return eng.axpby(1.0 / np.sqrt(2.0), Z, 0.0, nv.empty(int(Z.shape[0]), int(Z.shape[1])))
