"""Electronic circular dichroism (reference ``pymolresponse/properties/ecd.py:20-107``): rotational
strengths from the transition moments a TDHF / TDA driver left on the operators."""

from __future__ import annotations

from typing import Any

import numpy as np

from pymolresponse_b200 import integrals
from pymolresponse_b200.constants import HARTREE_TO_EV, HARTREE_TO_INVCM, alpha, esuecd
from pymolresponse_b200.molecular_property import TransitionProperty
from pymolresponse_b200.operators import Operator


class ECD(TransitionProperty):
    """Operator order: angular momentum (imaginary), dipole length (real), optionally dipole
    velocity (imaginary)."""

    def __init__(self, program, program_obj: Any, driver, *, do_dipvel: bool = False) -> None:
        super().__init__(program, program_obj, driver)
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
        ops = self.driver.solver.operators
        operator_angmom, operator_diplen = ops[0], ops[1]
        tm_l, tm_r = operator_angmom.transition_moments, operator_diplen.transition_moments
        assert len(tm_l) == len(tm_r)
        # R_n = esuecd * (-1/2) * <0|r|n> . <n|L|0>   (ecd.py:83-92)
        self.rotational_strengths_diplen = esuecd * (-1 / 2) * np.einsum("sc,sc->s", tm_r, tm_l)
        if self.do_dipvel:
            assert len(ops) == 3
            eigvals = np.asarray(self.driver.solver.eigvals).real
            tm_p = ops[2].transition_moments
            assert len(tm_p) == len(tm_r)
            self.rotational_strengths_dipvel = (
                esuecd * (-1 / 2) * np.einsum("sc,sc->s", tm_p / eigvals[:, None], tm_l))

    # ------------------------------------------------------------------ reports
    def print_results_nwchem(self) -> str:
        lines = [self.driver.print_results_nwchem()]
        energies = np.asarray(self.driver.solver.eigvals).real
        ops = self.driver.solver.operators
        tmom_angmom, tmom_diplen = ops[0].transition_moments, ops[1].transition_moments
        bar = "  " + "-" * 76
        for state in range(len(energies)):
            lines += [bar,
                      f"  Root {state + 1:>3d} singlet a{energies[state]:>25.9f} a.u."
                      f"{energies[state] * HARTREE_TO_EV:>22.4f} eV", bar,
                      "     Transition Moments    "
                      f"X{tmom_diplen[state, 0]:>9.5f}   Y{tmom_diplen[state, 1]:>9.5f}   "
                      f"Z{tmom_diplen[state, 2]:>9.5f}",
                      f"     Dipole Oscillator Strength {ops[1].total_oscillator_strengths[state]:>31.5f}",
                      "", "     Electric Transition Dipole:",
                      f"            X{tmom_diplen[state, 0]:>13.7f}   Y{tmom_diplen[state, 1]:>13.7f}   "
                      f"Z{tmom_diplen[state, 2]:>13.7f}",
                      "     Magnetic Transition Dipole (Length):",
                      f"            X{tmom_angmom[state, 0]:>13.7f}   Y{tmom_angmom[state, 1]:>13.7f}   "
                      f"Z{tmom_angmom[state, 2]:>13.7f}",
                      "     Magnetic Transition Dipole * 1/c :",
                      f"            X{tmom_angmom[state, 0] * alpha:>13.7f}   "
                      f"Y{tmom_angmom[state, 1] * alpha:>13.7f}   Z{tmom_angmom[state, 2] * alpha:>13.7f}",
                      "     Rotatory Strength (1E-40 esu**2cm**2):"
                      f"{self.rotational_strengths_diplen[state]:>21.7f}", ""]
            if self.do_dipvel:
                tmom_dipvel = ops[2].transition_moments
                lines += ["     Electric Transition Dipole (velocity representation):",
                          f"            X{tmom_dipvel[state, 0]:>13.7f}   Y{tmom_dipvel[state, 1]:>13.7f}   "
                          f"Z{tmom_dipvel[state, 2]:>13.7f}",
                          "     Oscillator Strength (velocity repr.) :"
                          f"{ops[2].total_oscillator_strengths[state]:>21.7f}",
                          "     Rotatory Strength   (velocity repr.) :"
                          f"{self.rotational_strengths_dipvel[state]:>21.7f}", ""]
        return "\n".join(lines)

    def print_results_orca(self) -> str:
        lines = [self.driver.print_results_orca()]
        energies = np.asarray(self.driver.solver.eigvals).real
        invcm = energies * HARTREE_TO_INVCM
        nm = 10000000 / invcm
        ops = self.driver.solver.operators

        def table(title, head, fosc, tm):
            t2 = np.einsum("sc,sc->s", tm, tm)
            out = ["-" * 77, title, "-" * 77, head[0], head[1], "-" * 77]
            out += [f"{s + 1:>4d}{invcm[s]:>10.1f}{nm[s]:>9.1f}{fosc[s]:>14.9f}{t2[s]:>10.5f}"
                    f"{tm[s, 0]:>10.5f}{tm[s, 1]:>10.5f}{tm[s, 2]:>10.5f}" for s in range(len(energies))]
            return out + [""]

        lines += table("         ABSORPTION SPECTRUM VIA TRANSITION ELECTRIC DIPOLE MOMENTS",
                       ("State   Energy  Wavelength   fosc         T2         TX        TY        TZ",
                        "        (cm-1)    (nm)                  (au**2)     (au)      (au)      (au)"),
                       ops[1].total_oscillator_strengths, ops[1].transition_moments)
        if self.do_dipvel:
            lines += table("         ABSORPTION SPECTRUM VIA TRANSITION VELOCITY DIPOLE MOMENTS",
                           ("State   Energy  Wavelength   fosc         P2         PX        PY        PZ",
                            "        (cm-1)    (nm)                  (au**2)     (au)      (au)      (au)"),
                           ops[2].total_oscillator_strengths, ops[2].transition_moments)
        tm_l = ops[0].transition_moments
        lines += ["-" * 67, "                             CD SPECTRUM", "-" * 67,
                  "State   Energy Wavelength       R         MX        MY        MZ",
                  "        (cm-1)   (nm)       (1e40*cgs)   (au)      (au)      (au)", "-" * 67]
        lines += [f"{s + 1:>4d}{invcm[s]:>10.1f}{nm[s]:>9.1f}{self.rotational_strengths_diplen[s]:>13.5f}"
                  f"{tm_l[s, 0]:>10.5f}{tm_l[s, 1]:>10.5f}{tm_l[s, 2]:>10.5f}" for s in range(len(energies))]
        lines.append("")
        return "\n".join(lines)
