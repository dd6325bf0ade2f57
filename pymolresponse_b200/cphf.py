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

    def _block_matrices(self, get_p, get_r, nfreq):
        """results[f][rows(op1), cols(op2)] = sum_sigma P_op1,sigma . R_op2,sigma[f]^T for every
        frequency f: ONE contraction per spin with all operators and frequencies stacked."""
        ops = self.solver.operators
        tags = ("alph", "beta") if self.solver.is_uhf else ("alph",)
        torch = nv.torch_mod()
        dims = [int(get_p(op, "alph").shape[0]) for op in ops]
        total = sum(dims)
        out = np.zeros((nfreq, total, total))
        if total == 0 or nfreq == 0:
            return out
        for tag in tags:
            P = torch.cat([get_p(op, tag) for op in ops], dim=0)
            R = torch.cat([get_r(op, tag, f) for f in range(nfreq) for op in ops], dim=0)
            full = nv.to_host(contract_rows(P, R))            # (total, nfreq * total)
            out += full.reshape(total, nfreq, total).transpose(1, 0, 2)
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
        freqs = [float(w) for w in solver.frequencies]

        def divided(op, tag, f):
            sv = op.supervector_dev(tag)
            ncomp, two_nov = int(sv.shape[0]), int(sv.shape[1])
            out = nv.empty(ncomp, two_nov)
            nv.check(L.pmr_uncoupled_divide(nv.ptr(sv), ncomp, two_nov // 2, nv.ptr(ediffs[tag]),
                                            freqs[f], nv.ptr(out), nv.stream_ptr()))
            return out

        allres = self._block_matrices(lambda op, tag: op.supervector_dev(tag), divided, len(freqs))
        for f in range(len(freqs)):
            results = allres[f]
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
        nfreq = len(solver.frequencies)
        allres = self._block_matrices(lambda op, tag: op.supervector_dev(tag),
                                      lambda op, tag, f: self._rows(op, "rsp", tag, f), nfreq)
        for f in range(nfreq):
            results = allres[f]
            if solver.is_uhf:
                results = 2 * results / 2
            else:
                results = 4 * results / 2
            self.results.append(results)
