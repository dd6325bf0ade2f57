"""Driver for the coupled-perturbed Hartree-Fock (CPHF) equations.

Drop-in for the reference's ``pymolresponse/cphf.py`` (``Driver``, ``CPHF``): same methods, same
``results`` / ``uncoupled_results`` lists of ``(sum ncomp, sum ncomp)`` float64 arrays.  The
contractions ``pref * sum_sigma P R^T`` (cphf.py:165-253, utils.py:17-27) and the uncoupled
``P (P / (de -+ w))^T`` (cphf.py:68-163) run on the GPU over the device-resident vectors the
solver left behind; only the small result matrices come back to the host.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Sequence
from typing import Any

import numpy as np

from pymolresponse_b200 import _native as nv
from pymolresponse_b200.utils import contract_rows


class Driver(ABC):
    def __init__(self, solver) -> None:
        self.solver = solver
        self.results = []

    @abstractmethod
    def run(self, hamiltonian, spin, program, program_obj: Any) -> None:
        pass


class CPHF(Driver):
    """Driver for solving the CPHF equations (reference cphf.py:31-253)."""

    def __init__(self, solver) -> None:
        # as in the reference, Driver.__init__ is not called: ``results`` exists after ``run``
        self.solver = solver

    def set_frequencies(self, frequencies: Sequence[float] | None = None) -> None:
        r"""Set the frequencies :math:`\omega_f` (atomic units); ``None`` means static."""
        assert self.solver is not None
        self.solver.set_frequencies(frequencies)
        self.frequencies = self.solver.frequencies

    def add_operator(self, operator) -> None:
        """Add a right-hand-side perturbation."""
        assert self.solver is not None
        self.solver.add_operator(operator)

    def run(self, hamiltonian, spin, program, program_obj: Any) -> None:
        if not hasattr(self, "frequencies"):
            self.set_frequencies([0.0])
        self.solver.run(hamiltonian, spin, program, program_obj)
        self.form_uncoupled_results()
        self.form_results()

    # ------------------------------------------------------------------ helpers
    def _rows(self, op, what: str, tag: str, f: int | None = None):
        """Device (ncomp, 2nov) tensor of an operator's supervectors / response vectors."""
        if what == "super":
            return op.supervector_dev(tag)
        dev = getattr(op, f"_rspvecs_{tag}_dev")
        host = getattr(op, f"rspvecs_{tag}")
        if f < len(dev) and len(dev) == len(host):
            return dev[f]
        return nv.to_dev(host[f])[:, :, 0].contiguous()

    def _block_matrix(self, get_p, get_r):
        """results[rows(op1), cols(op2)] = sum_sigma P_op1,sigma . R_op2,sigma^T."""
        ops = self.solver.operators
        tags = ("alph", "beta") if self.solver.is_uhf else ("alph",)
        dims = [int(get_p(op, "alph").shape[0]) for op in ops]
        total = sum(dims)
        out = np.zeros((total, total))
        # stack all operators: one contraction per spin gives every block at once
        for tag in tags:
            P = nv.torch_mod().cat([get_p(op, tag) for op in ops], dim=0)
            R = nv.torch_mod().cat([get_r(op, tag) for op in ops], dim=0)
            out += nv.to_host(contract_rows(P, R))
        return out

    def form_uncoupled_results(self) -> None:
        """Uncoupled (sum-over-orbital-pairs) results (reference cphf.py:68-163)."""
        solver = self.solver
        nocc_a, _, nocc_b, _ = solver.occupations
        L = nv.lib()
        ediffs = {}
        for tag, spin, nocc in (("alph", 0, nocc_a),) + ((("beta", 1, nocc_b),) if solver.is_uhf else ()):
            eps = nv.to_dev(solver.moenergies[spin]).diagonal().contiguous()
            eo, ev = eps[:nocc].contiguous(), eps[nocc:].contiguous()
            ed = nv.empty(max(1, int(eo.shape[0]) * int(ev.shape[0])))
            nv.check(L.pmr_energy_differences(nv.ptr(eo), int(eo.shape[0]), nv.ptr(ev),
                                              int(ev.shape[0]), nv.ptr(ed), nv.stream_ptr()))
            ediffs[tag] = ed
        self.uncoupled_results = []
        for f in range(len(solver.frequencies)):
            frequency = float(solver.frequencies[f])

            def divided(op, tag):
                sv = op.supervector_dev(tag)
                ncomp, two_nov = int(sv.shape[0]), int(sv.shape[1])
                out = nv.empty(ncomp, two_nov)
                nv.check(L.pmr_uncoupled_divide(nv.ptr(sv), ncomp, two_nov // 2,
                                                nv.ptr(ediffs[tag]), frequency, nv.ptr(out),
                                                nv.stream_ptr()))
                return out

            results = self._block_matrix(lambda op, tag: op.supervector_dev(tag), divided)
            # the / 2 is because of the supervector part (reference cphf.py:158-162)
            if solver.is_uhf:
                results = 2 * results / 2
            else:
                results = 4 * results / 2
            self.uncoupled_results.append(results)

    def form_results(self) -> None:
        """Coupled results from the response vectors (reference cphf.py:165-253)."""
        solver = self.solver
        self.results = []
        for operator in solver.operators:
            assert len(solver.frequencies) == len(operator.rspvecs_alph)
            if solver.is_uhf:
                assert len(solver.frequencies) == len(operator.rspvecs_beta)
            else:
                assert len(operator.rspvecs_beta) == 0
        for f in range(len(solver.frequencies)):
            for op in solver.operators:
                assert tuple(op.rspvecs_alph[f].shape) == tuple(
                    op.mo_integrals_ai_supervector_alph.shape)
                if solver.is_uhf:
                    assert tuple(op.rspvecs_beta[f].shape) == tuple(
                        op.mo_integrals_ai_supervector_beta.shape)
        for f in range(len(solver.frequencies)):
            results = self._block_matrix(
                lambda op, tag: op.supervector_dev(tag),
                lambda op, tag, f=f: self._rows(op, "rsp", tag, f),
            )
            if solver.is_uhf:
                results = 2 * results / 2
            else:
                results = 4 * results / 2
            self.results.append(results)
