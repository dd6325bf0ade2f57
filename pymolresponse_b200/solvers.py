"""Linear-response solvers on the GPU.

Drop-in for the linear-equation part of the reference's ``pymolresponse/solvers.py``
(``Solver`` -> ``LineqSolver`` -> ``ExactLineqSolver`` -> ``ExactInv`` / ``ExactInvCholesky`` and
``IterativeLinEqSolver``): same constructors, same attributes (``tei_mo``, ``tei_mo_type``,
``explicit_hessian``, ``explicit_hessian_inv``, ``left_alph``/``left_beta``, ``operators``,
``frequencies``), same side effects (``operator.rspvecs_alph/beta.append(ndarray(ncomp, 2nov, 1))``
once per frequency) and the same exceptions.  What differs is how the numbers are produced:

* the A/B blocks are assembled ONCE per run by ``pmr_assemble`` directly as M = A - B and
  K = A + B (the reference re-forms A and B at every frequency, solvers.py:477-478);
* no (2 nov)^2 matrix is built, inverted or multiplied: see ``_engine.HalfSizeSystem``;
  ``explicit_hessian`` / ``explicit_hessian_inv`` are materialised lazily, on access, for the
  last frequency (parity checks and downstream code that reads them);
* UHF uses the joint (nov_a + nov_b) system instead of four explicit inverses and two Schur
  complements (solvers.py:506-520).
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Callable, Sequence
from typing import Any

import numpy as np

from pymolresponse_b200 import _blocks as bl
from pymolresponse_b200 import _engine as eng
from pymolresponse_b200 import _native as nv
from pymolresponse_b200.core import AO2MOTransformationType, Hamiltonian, Program, Spin
from pymolresponse_b200.indices import form_indices_from_occupations
from pymolresponse_b200.integrals import JK
from pymolresponse_b200.operators import Operator


class Solver(ABC):
    """State shared by all solvers (reference solvers.py:41-154)."""

    def __init__(self, mocoeffs, moenergies, occupations) -> None:
        # MO coefficients: first axis is alpha/beta
        assert len(mocoeffs.shape) == 3
        assert (mocoeffs.shape[0] == 1) or (mocoeffs.shape[0] == 2)
        self.is_uhf = mocoeffs.shape[0] == 2
        # MO energies: first axis alpha/beta, then a diagonal square matrix
        assert len(moenergies.shape) == 3
        assert (moenergies.shape[0] == 1) or (moenergies.shape[0] == 2)
        if self.is_uhf:
            assert moenergies.shape[0] == 2
        else:
            assert moenergies.shape[0] == 1
        assert moenergies.shape[1] == moenergies.shape[2]
        assert len(occupations) == 4

        self.mocoeffs = mocoeffs
        self.moenergies = moenergies
        self.occupations = occupations
        self.indices = form_indices_from_occupations(occupations)

        self.operators = []
        self.frequencies = []

        self._tei_mo = None
        self._tei_mo_dev = None
        self.tei_mo_type = AO2MOTransformationType.partial
        self._explicit_hessian = None
        self._explicit_hessian_inv = None
        self._lazy_hessian = None
        # keep response vectors on the device as torch tensors instead of NumPy arrays
        self.device_results = False
        # "s8": form_tei_mo may assume the 8-fold symmetry of the AO ERI tensor (see ao2mo.py)
        self.ao_symmetry = "s1"
        # row-block-sharded Hessian: ``tei_mo`` holds only this rank's occupied rows of every block
        # (AO2MO(reduce="scatter").row_ranges): (lo, hi) for RHF, ((lo, hi), (lo, hi)) for UHF
        self.tei_mo_rows = None

    # ---------------------------------------------------------------- MO integrals
    explicit_hessian = None
    explicit_hessian_inv = None

    @property
    def tei_mo(self):
        """Tuple of MO-basis ERI blocks (length 1/2/4/6, reference solvers.py:170)."""
        return self._tei_mo

    @tei_mo.setter
    def tei_mo(self, value):
        self._tei_mo = value
        self._tei_mo_dev = None  # device mirror is rebuilt on demand

    def form_tei_mo(self, program: Program, program_obj: Any,
                    tei_mo_type: AO2MOTransformationType) -> None:
        """AO -> MO transformation on the GPU (reference solvers.py:86-127).

        ``program_obj`` may be the AO ERI tensor itself (NumPy / CUDA, ``Program.Arrays``), an
        object with an ``ao_eri`` attribute, a pyscf ``Mole`` (``Program.PySCF``) or a psi4
        molecule (``Program.Psi4``); the quantum-chemistry packages are imported lazily and only
        to *generate* the AO integrals.
        """
        from pymolresponse_b200.ao2mo import AO2MO

        nden = self.mocoeffs.shape[0]
        assert nden in (1, 2)
        I = None  # noqa: E741
        df_factor = getattr(program_obj, "df_factor", None)
        if df_factor is not None:
            pass  # factorised AO integrals: no n^4 tensor (ao2mo.py, df_factor)
        elif isinstance(program_obj, np.ndarray) or nv.is_device_array(program_obj):
            I = program_obj  # noqa: E741
        elif hasattr(program_obj, "ao_eri"):
            I = program_obj.ao_eri  # noqa: E741
        elif program == Program.PySCF:
            I = program_obj.intor("int2e", aosym="s1")  # noqa: E741
        elif program == Program.Psi4:
            import psi4

            I = np.asarray(  # noqa: E741
                psi4.core.MintsHelper(
                    psi4.core.Wavefunction.build(
                        program_obj, psi4.core.get_global_option("BASIS")
                    ).basisset()
                ).ao_eri()
            )
        else:
            raise RuntimeError
        ao2mo = AO2MO(self.mocoeffs, self.occupations, I=I, device_results=True,
                      ao_symmetry=self.ao_symmetry, df_factor=df_factor)
        if tei_mo_type == AO2MOTransformationType.partial and nden == 2:
            ao2mo.perform_uhf_partial()
        elif tei_mo_type == AO2MOTransformationType.partial and nden == 1:
            ao2mo.perform_rhf_partial()
        elif tei_mo_type == AO2MOTransformationType.full and nden == 2:
            ao2mo.perform_uhf_full()
        elif tei_mo_type == AO2MOTransformationType.full and nden == 1:
            ao2mo.perform_rhf_full()
        else:
            msg = f"Unknown combination of tei_mo_type {tei_mo_type} and nden {nden}"
            raise ValueError(msg)
        dev = tuple(ao2mo.tei_mo)
        self.tei_mo = dev if self.device_results else tuple(nv.to_host(t) for t in dev)
        self._tei_mo_dev = list(dev)
        self.tei_mo_type = tei_mo_type

    def set_frequencies(self, frequencies: Sequence[float] | None = None) -> None:
        if frequencies is None:
            self.frequencies = [0.0]
        else:
            self.frequencies = frequencies
        if hasattr(self, "operators"):
            for operator in self.operators:
                operator.frequencies = self.frequencies

    def add_operator(self, operator: Operator) -> None:
        # first dimension: Cartesian components; then (nao, nao)
        assert hasattr(operator, "ao_integrals")
        shape = operator.ao_integrals.shape
        assert len(shape) == 3
        assert shape[0] >= 1
        assert shape[1] == shape[2]
        operator.form_rhs(self.mocoeffs, self.occupations, self.indices)
        self.operators.append(operator)

    @abstractmethod
    def run(self, hamiltonian: Hamiltonian, spin: Spin, program: Program | None,
            program_obj: Any) -> None:
        """Run the solver."""

    # ---------------------------------------------------------------- helpers
    def _append_rspvecs(self, operator, tag, xy):
        """xy: (ncomp, 2nov) CUDA tensor for one frequency."""
        getattr(operator, f"_rspvecs_{tag}_dev").append(xy)
        if self.device_results:
            getattr(operator, f"rspvecs_{tag}").append(xy[:, :, None])
        else:
            getattr(operator, f"rspvecs_{tag}").append(nv.to_host(xy)[:, :, None])


class LineqSolver(Solver, ABC):
    """Base class for all solvers of the type AX = B."""


def _kind(spin: Spin, uhf: bool) -> str:
    if spin == Spin.triplet:
        return "triplet"
    return "ss" if uhf else "singlet"


class _HessianBlocks:
    """Device-side assembly of the orbital-Hessian blocks from the MO integrals, shared by the
    linear-equation and the eigenvalue solvers (reference solvers.py:166-385 and :699-752)."""

    def _tei_dev(self):
        assert self.tei_mo is not None
        assert len(self.tei_mo) in (1, 2, 4, 6)
        assert isinstance(self.tei_mo_type, AO2MOTransformationType)
        if self._tei_mo_dev is None:
            self._tei_mo_dev = [nv.to_dev(t) for t in self.tei_mo]
        return self._tei_mo_dev

    def _assemble_rhf(self, hamiltonian, spin, want_ab, want_mk=True):
        """-> (M, K, A, B); A/B only when want_ab (LU fallback, explicit_hessian)."""
        nocc, nvirt, _, _ = self.occupations
        nov = nocc * nvirt
        tei = self._tei_dev()
        kind = _kind(spin, False)
        eps = bl.split_energies(self.moenergies[0], nocc)
        rows = self._row_ranges()
        if rows is not None:
            return self._assemble_rhf_rows(hamiltonian, kind, eps, rows[0], want_ab, want_mk)
        if self.tei_mo_type == AO2MOTransformationType.full:
            assert len(tei) == 1
            terms = bl.full_terms(kind, tei[0], nocc, nocc)
        else:
            assert len(tei) == 2
            terms = bl.partial_terms(kind, tei[0], tei[1])
        tda = hamiltonian == Hamiltonian.TDA
        if tda:
            terms = (terms[0], terms[1], bl.absent(), bl.absent())
        A = B = M = K = None
        if tda:
            A, _ = bl.run(nocc, nvirt, nocc, nvirt, terms, eps, "P")
            M = K = A
        else:
            if want_mk:
                M = nv.empty(nov, nov)
                K = nv.empty(nov, nov)
                bl.run(nocc, nvirt, nocc, nvirt, terms, eps, "PQ", out_mode=1, out0=M, out1=K, ld=nov)
            if want_ab:
                A, B = bl.run(nocc, nvirt, nocc, nvirt, terms, eps, "PQ")
        return M, K, A, B

    def _row_ranges(self):
        from pymolresponse_b200 import distributed as pd

        rows = getattr(self, "tei_mo_rows", None)
        if rows is None or pd.rank_world()[1] == 1:
            return None
        if isinstance(rows[0], int):
            rows = (tuple(rows), tuple(rows))
        return rows

    @staticmethod
    def _gather_row_blocks(loc, ri, nv_x, ncols, no_total):
        """Row block (ri * nv_x, ncols) of every rank -> the whole (no_total * nv_x, ncols) matrix."""
        from pymolresponse_b200 import distributed as pd

        return pd.allgather_rows(loc.view(ri, nv_x * ncols), no_total).view(no_total * nv_x, ncols)

    def _assemble_rhf_rows(self, hamiltonian, kind, eps, rows, want_ab, want_mk):
        """Row-block-sharded assembly (north_star: 'the ov Hessian over row blocks'): this rank forms
        the rows of its occupied indices [i_lo, i_hi) from its rows of (ia|jb) and (ij|ab) -- both
        index swaps of the A/B formulas stay inside a row block -- and the row blocks are all-gathered
        because every rank factorises its own frequencies."""
        nocc, nvirt, _, _ = self.occupations
        nov = nocc * nvirt
        i_lo, i_hi = rows
        ri = i_hi - i_lo
        assert self.tei_mo_type == AO2MOTransformationType.partial, "row sharding: partial blocks only"
        tei = self._tei_dev()
        assert tuple(tei[0].shape) == (ri, nvirt, nocc, nvirt) and tuple(tei[1].shape) == (ri, nocc, nvirt, nvirt)
        terms = bl.partial_terms(kind, tei[0], tei[1], i_lo)
        tda = hamiltonian == Hamiltonian.TDA
        if tda:
            terms = (terms[0], terms[1], bl.absent(), bl.absent())
        A = B = M = K = None

        def gather(t):
            return self._gather_row_blocks(t, ri, nvirt, nov, nocc)

        if tda:
            A_loc, _ = bl.run(nocc, nvirt, nocc, nvirt, terms, eps, "P", i_lo=i_lo, i_hi=i_hi)
            A = gather(A_loc)
            M = K = A
        else:
            if want_mk:
                M_loc, K_loc = nv.empty(ri * nvirt, nov), nv.empty(ri * nvirt, nov)
                bl.run(nocc, nvirt, nocc, nvirt, terms, eps, "PQ", out_mode=1, out0=M_loc, out1=K_loc,
                       ld=nov, i_lo=i_lo, i_hi=i_hi)
                M, K = gather(M_loc), gather(K_loc)
            if want_ab:
                A_loc, B_loc = bl.run(nocc, nvirt, nocc, nvirt, terms, eps, "PQ", i_lo=i_lo, i_hi=i_hi)
                A, B = gather(A_loc), gather(B_loc)
        return M, K, A, B

    def _assemble_uhf(self, hamiltonian, spin, want_ab, want_mk=True):
        """Joint (nov_a + nov_b) matrices [[ss_a, os_a],[os_b, ss_b]] -> (M, K, A, B)."""
        na, va, nb, vb = self.occupations
        nova, novb = na * va, nb * vb
        nt = nova + novb
        tei = self._tei_dev()
        full = self.tei_mo_type == AO2MOTransformationType.full
        assert len(tei) == (4 if full else 6)
        tda = hamiltonian == Hamiltonian.TDA
        kind = _kind(spin, True)
        eps_a = bl.split_energies(self.moenergies[0], na)
        eps_b = bl.split_energies(self.moenergies[1], nb)

        def terms(block, i_lo=0):
            if full:
                aaaa, aabb, bbaa, bbbb = tei
                t = {"aa": (kind, aaaa, na, na), "bb": (kind, bbbb, nb, nb),
                     "ab": ("os", aabb, na, nb), "ba": ("os", bbaa, nb, na)}[block]
                out = bl.full_terms(*t)
            else:
                ovov_aaaa, ovov_aabb, ovov_bbaa, ovov_bbbb, oovv_aaaa, oovv_bbbb = tei
                t = {"aa": (kind, ovov_aaaa, oovv_aaaa), "bb": (kind, ovov_bbbb, oovv_bbbb),
                     "ab": ("os", ovov_aabb, None), "ba": ("os", ovov_bbaa, None)}[block]
                out = bl.partial_terms(*t, i_lo)
            if tda:
                out = (out[0], out[1], bl.absent(), bl.absent())
            return out

        rows = self._row_ranges()
        if rows is not None:
            assert not full, "row sharding: partial blocks only"
            (ia_lo, ia_hi), (ib_lo, ib_hi) = rows
        else:
            (ia_lo, ia_hi), (ib_lo, ib_hi) = (0, na), (0, nb)

        def fill(out0, out1, mode):
            # offsets of the four blocks inside an (nt x nt) matrix
            blocks = (("aa", na, va, na, va, eps_a, 0), ("bb", nb, vb, nb, vb, eps_b, nova * nt + nova),
                      ("ab", na, va, nb, vb, None, nova), ("ba", nb, vb, na, va, None, nova * nt))
            for name, ox, vx, oy, vy, eps, off in blocks:
                i_lo, i_hi = (ia_lo, ia_hi) if name[0] == "a" else (ib_lo, ib_hi)
                if name in ("ab", "ba") and spin == Spin.triplet:
                    # no Coulomb term, hence no opposite-spin coupling (solvers.py:272-285)
                    cols = oy * vy
                    r0, c0 = divmod(off, nt)
                    for o in (out0, out1):
                        if o is not None:
                            o[r0 + i_lo * vx:r0 + i_hi * vx, c0:c0 + cols].zero_()
                    continue
                # a sharded rank fills only its occupied rows [i_lo, i_hi) of the (nt x nt) matrix
                # (global row addressing, so the offsets are those of the full matrix)
                bl.run(ox, vx, oy, vy, terms(name, i_lo if rows is not None else 0), eps, "PQ",
                       out_mode=mode, out0=out0, out1=out1, ld=nt,
                       off0=off + i_lo * vx * nt, off1=off + i_lo * vx * nt, i_lo=i_lo, i_hi=i_hi)
            if rows is not None:
                # exchange the row blocks: alpha rows and beta rows of every output
                for o in (out0, out1):
                    if o is None:
                        continue
                    for r0, vx, ox, (lo, hi) in ((0, va, na, (ia_lo, ia_hi)), (nova, vb, nb, (ib_lo, ib_hi))):
                        loc = o[r0 + lo * vx:r0 + hi * vx].contiguous()
                        o[r0:r0 + ox * vx].copy_(self._gather_row_blocks(loc, hi - lo, vx, nt, ox))

        A = B = M = K = None
        if tda:
            A = nv.empty(nt, nt)
            fill(A, None, 0)
            M = K = A
        else:
            if want_mk:
                M, K = nv.empty(nt, nt), nv.empty(nt, nt)
                fill(M, K, 1)
            if want_ab:
                A, B = nv.empty(nt, nt), nv.empty(nt, nt)
                fill(A, B, 0)
        return M, K, A, B


class ExactLineqSolver(LineqSolver, _HessianBlocks, ABC):
    """Solvers that (conceptually) form the orbital Hessian explicitly
    (reference solvers.py:160-480)."""

    #: ExactInv falls back to pivoted LU on G(w) when a Cholesky pivot fails; the Cholesky
    #: solver raises instead (the reference's scipy.linalg.cholesky would).
    _allow_lu_fallback = True

    def share_hessian_with(self, other: "ExactLineqSolver") -> None:
        """Let ``other`` reuse this solver's assembled M / K matrices, the Cholesky factor of K and
        K^-1 (all device resident) instead of rebuilding them: several properties computed for the
        same reference wavefunction, Hamiltonian and spin need them only once.  Call it after this
        solver has run (or at least been prepared); ``other`` must use the same MO integrals."""
        assert getattr(self, "_system", None) is not None, "run or prepare this solver first"
        assert self.occupations == other.occupations and self.is_uhf == other.is_uhf
        other._system = self._system
        other._prepared_for = self._prepared_for
        other._hessian_shared = True

    def _prepare(self, hamiltonian: Hamiltonian, spin: Spin, want_ab: bool = False):
        assemble = self._assemble_uhf if self.is_uhf else self._assemble_rhf
        with nv.nvtx_range("assemble_hessian"):
            M, K, A, B = assemble(hamiltonian, spin, want_ab)
        self._system = eng.HalfSizeSystem(M, K, A, B)
        self._prepared_for = (hamiltonian, spin)
        return self._system

    # -------------------------------------------------------------- reference API
    def form_explicit_hessian(self, hamiltonian: Hamiltonian, spin: Spin,
                              frequency: float | None) -> None:
        """Materialise G(w) exactly as the reference lays it out (solvers.py:166-385): an
        ndarray (2nov, 2nov) for RHF, the 4-tuple (G_aa - wD, G_ab, G_ba, G_bb - wD) for UHF."""
        assert self.tei_mo is not None
        assert len(self.tei_mo) in (1, 2, 4, 6)
        assert isinstance(self.tei_mo_type, AO2MOTransformationType)
        assert isinstance(frequency, float)
        self._last_call = (hamiltonian, spin, frequency)
        _, _, A, B = (self._assemble_uhf if self.is_uhf else self._assemble_rhf)(
            hamiltonian, spin, True, want_mk=False)
        if hamiltonian == Hamiltonian.TDA:
            B = None
        if not self.is_uhf:
            self._explicit_hessian = nv.to_host(eng.super_hessian(A, B, frequency))
        else:
            na, va, nb, vb = self.occupations
            nova = na * va

            def sup(r, c, w):
                Ab = A[r, c].contiguous()
                Bb = B[r, c].contiguous() if B is not None else None
                if w is None:
                    N1, N2 = Ab.shape
                    G = nv.zeros(2 * N1, 2 * N2)
                    G[:N1, :N2] = Ab
                    G[N1:, N2:] = Ab
                    if Bb is not None:
                        G[:N1, N2:] = Bb
                        G[N1:, :N2] = Bb
                    return nv.to_host(G)
                return nv.to_host(eng.super_hessian(Ab, Bb, w))

            a, b = slice(0, nova), slice(nova, None)
            self._explicit_hessian = (sup(a, a, frequency), sup(a, b, None), sup(b, a, None),
                                      sup(b, b, frequency))
        self._lazy_hessian = None

    @property
    def explicit_hessian(self):
        if self._explicit_hessian is None and self._lazy_hessian is not None:
            ham, spin, w = self._lazy_hessian
            self.form_explicit_hessian(ham, spin, w)
        return self._explicit_hessian

    @explicit_hessian.setter
    def explicit_hessian(self, value):
        self._explicit_hessian = value
        self._lazy_hessian = None

    @staticmethod
    def _gpu_inverse(G: np.ndarray) -> np.ndarray:
        Gd = nv.to_dev(G).clone()
        ipiv, info = eng.getrf(Gd)
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError("Singular matrix")
        N = int(Gd.shape[0])
        eye = nv.zeros(N, N)
        eng.shift_diagonal(eye, N, 1.0)
        return nv.to_host(eng.getrs(Gd, ipiv, eye))

    @staticmethod
    def _gpu_inverse_cholesky(G: np.ndarray) -> np.ndarray:
        """G^-1 through the Cholesky factor, raising like ``scipy.linalg.cholesky`` when G is not
        positive definite (reference solvers.py:527-530)."""
        Gd = nv.to_dev(G).clone()
        dinv, info = eng.potrf(Gd)
        if int(info.item()) != 0:
            raise np.linalg.LinAlgError(
                f"{int(info.item())}-th leading minor of the array is not positive definite")
        low = eng.potri_lower(Gd, dinv).tril()
        return nv.to_host(low + low.tril(-1).t())

    @property
    def explicit_hessian_inv(self):
        if self._explicit_hessian_inv is None and self.explicit_hessian is not None:
            self.invert_explicit_hessian()
        return self._explicit_hessian_inv

    # UHF Schur-complement inverses (reference solvers.py:519-520): the reference sets them in
    # ``invert_explicit_hessian`` on every run; here they are materialised on first access
    _left_alph = None
    _left_beta = None

    @property
    def left_alph(self):
        if self._left_alph is None and self.is_uhf and self.explicit_hessian is not None:
            self.invert_explicit_hessian()
        if self._left_alph is None:
            raise AttributeError("left_alph")
        return self._left_alph

    @left_alph.setter
    def left_alph(self, value):
        self._left_alph = value

    @property
    def left_beta(self):
        if self._left_beta is None and self.is_uhf and self.explicit_hessian is not None:
            self.invert_explicit_hessian()
        if self._left_beta is None:
            raise AttributeError("left_beta")
        return self._left_beta

    @left_beta.setter
    def left_beta(self, value):
        self._left_beta = value

    @explicit_hessian_inv.setter
    def explicit_hessian_inv(self, value):
        self._explicit_hessian_inv = value

    @abstractmethod
    def invert_explicit_hessian(self) -> None:
        """Invert the explicitly constructed electronic Hessian (parity/inspection only: the
        response vectors never use an explicit inverse)."""

    def _invert_with(self, inv) -> None:
        G = self.explicit_hessian
        if not self.is_uhf:
            assert isinstance(G, np.ndarray)
            assert len(G.shape) == 2 and G.shape[0] == G.shape[1]
            self._explicit_hessian_inv = inv(G)
        else:
            assert isinstance(G, tuple) and len(G) == 4
            G_aa, G_ab, G_ba, G_bb = G
            G_aa_inv, G_bb_inv = inv(G_aa), inv(G_bb)
            self._explicit_hessian_inv = [G_aa_inv, G_bb_inv]
            # Schur complements of the spin blocks (reference solvers.py:519-520), via the joint
            # inverse: left_alph = [G_joint^-1]_aa, left_beta = [G_joint^-1]_bb
            joint = np.block([[G_aa, G_ab], [G_ba, G_bb]])
            joint_inv = inv(joint)
            na2 = G_aa.shape[0]
            self.left_alph = np.ascontiguousarray(joint_inv[:na2, :na2])
            self.left_beta = np.ascontiguousarray(joint_inv[na2:, na2:])

    def form_response_vectors(self) -> None:
        """Solve for the frequency the last ``form_explicit_hessian`` call was made for and append
        to every operator (reference solvers.py:391-466)."""
        assert self._last_call is not None, "call form_explicit_hessian / run first"
        ham, spin, w = self._last_call
        self._solve_frequencies(ham, spin, [w])

    def _gradient_rows(self):
        """All operators' property gradients g stacked as rows, plus their supervector signs and a
        plan [(first row, ncomp, custom)] for :meth:`_store_xy`.  A supervector the caller replaced
        after ``form_rhs`` is a general [u; l]: it is split into a real-type column (u - l)/2 with
        sign -1 and an imaginary-type column (u + l)/2 with sign +1 (the equations are linear)."""
        na, va, nb, vb = self.occupations
        nova, novb = na * va, nb * vb
        torch = nv.torch_mod()
        rows, signs, plan = [], [], []
        for op in self.operators:
            tags = ("alph", "beta") if self.is_uhf else ("alph",)
            halves = []
            for tag, nov in zip(tags, (nova, novb)):
                sv = op.supervector_dev(tag)
                halves.append((sv[:, :nov], sv[:, nov:]))
            upper = torch.cat([h[0] for h in halves], dim=1) if self.is_uhf else halves[0][0]
            ncomp = int(upper.shape[0])
            custom = not all(op.has_standard_supervector(tag) for tag in tags)
            plan.append((len(signs), ncomp, custom))
            if not custom:
                rows.append(upper)
                signs += [+1 if op.is_imaginary else -1] * ncomp
            else:
                lower = torch.cat([h[1] for h in halves], dim=1) if self.is_uhf else halves[0][1]
                rows.append(0.5 * (upper - lower))
                rows.append(0.5 * (upper + lower))
                signs += [-1] * ncomp + [+1] * ncomp
        if not rows:
            return None, signs, plan
        return torch.cat(rows, dim=0).contiguous(), signs, plan

    def _solve_xy(self, hamiltonian, spin, frequencies, distributed=False, global_frequencies=None):
        """(nf, ncols, 2N) CUDA tensor of [X;Y] for the given frequencies (this rank's share)."""
        if getattr(self, "_prepared_for", None) != (hamiltonian, spin):
            self._prepare(hamiltonian, spin, want_ab=False)
        system = self._system
        g_rows, signs, plan = self._gradient_rows()
        self._plan = plan
        if g_rows is None:
            return None
        xy, bad = system.solve(g_rows, signs, frequencies, distributed=distributed,
                               global_frequencies=global_frequencies)
        if bad:
            freqs = [float(w) for w in frequencies]
            if not self._allow_lu_fallback:
                self.solve_stats = dict(system.stats)
                raise eng.LinAlgError(
                    "orbital Hessian is not positive definite at frequencies "
                    f"{[freqs[f] for f in sorted(bad)]} (Cholesky failed)")
            # pivoted LU on the full G(w) for the failed frequencies only; it needs the A and B
            # blocks themselves, assembled now (M, K and their factors stay as they are)
            if system.A is None:
                assemble = self._assemble_uhf if self.is_uhf else self._assemble_rhf
                _, _, system.A, system.B = assemble(hamiltonian, spin, True, want_mk=False)
            for f in sorted(bad):
                xy[f].copy_(system.solve_lu(g_rows, signs, freqs[f]))
                system.stats["lu_fallback"] += 1
        self.solve_stats = dict(system.stats)
        return xy

    def _store_xy(self, xy, n_frequencies):
        na, va, nb, vb = self.occupations
        nova, novb = na * va, nb * vb
        nt = nova + novb if self.is_uhf else nova
        torch = nv.torch_mod()
        for f in range(n_frequencies):
            for op, (c0, nc, custom) in zip(self.operators, self._plan):
                blk = xy[f][c0:c0 + nc]
                if custom:
                    blk = blk + xy[f][c0 + nc:c0 + 2 * nc]
                if self.is_uhf:
                    X, Y = blk[:, :nt], blk[:, nt:]
                    self._append_rspvecs(op, "alph",
                                         torch.cat((X[:, :nova], Y[:, :nova]), dim=1).contiguous())
                    self._append_rspvecs(op, "beta",
                                         torch.cat((X[:, nova:], Y[:, nova:]), dim=1).contiguous())
                else:
                    self._append_rspvecs(op, "alph", blk.contiguous())

    def _solve_frequencies(self, hamiltonian, spin, frequencies):
        """Solve every frequency (dealt round-robin to the ranks when torch.distributed is up and
        ``distribute_frequencies`` is set) and append the response vectors in frequency order."""
        from pymolresponse_b200 import distributed as pd

        rank, world = pd.rank_world()
        if world > 1 and getattr(self, "distribute_frequencies", False):
            mine = pd.round_robin(len(frequencies), rank, world)
            xy_local, err = None, None
            try:
                xy_local = self._solve_xy(hamiltonian, spin, [frequencies[i] for i in mine],
                                          distributed=True, global_frequencies=list(frequencies))
            except np.linalg.LinAlgError as e:
                err = e
            # agree on failure BEFORE the gather: a rank that raised alone would leave the others
            # waiting in the collective (the reference raises cleanly, solvers.py:527)
            if pd.any_rank(err is not None):
                raise err if err is not None else eng.LinAlgError(
                    "orbital Hessian is not positive definite at a frequency owned by another rank "
                    "(Cholesky failed)")
            if not self.operators:
                return
            xy = pd.gather_round_robin([xy_local[k] for k in range(len(mine))], len(frequencies))
        else:
            xy = self._solve_xy(hamiltonian, spin, frequencies)
            if xy is None:
                return
        self._store_xy(xy, len(frequencies))

    def run(self, hamiltonian: Hamiltonian, spin: Spin, program: Program | None,
            program_obj: Any) -> None:
        if not self.tei_mo:
            if program is None:
                msg = "Program must be specified for computing 2-electron integrals in MO basis"
                raise RuntimeError(msg)
            self.form_tei_mo(program, program_obj, AO2MOTransformationType.partial)
        for frequency in self.frequencies:
            assert isinstance(frequency, float)
        freqs = [float(w) for w in self.frequencies]
        if not getattr(self, "_hessian_shared", False):
            self._prepared_for = None
        if getattr(self, "inv_func", None) is not None and self.inv_func is not np.linalg.inv:
            # caller-injected inversion routine (reference solvers.py:490-497): honour it literally
            for w in freqs:
                self.form_explicit_hessian(hamiltonian, spin, w)
                self.invert_explicit_hessian()
                self._response_from_explicit_inverse()
            return
        if freqs:
            self._solve_frequencies(hamiltonian, spin, freqs)
            self._explicit_hessian = None
            self._explicit_hessian_inv = None
            self._left_alph = self._left_beta = None
            self._lazy_hessian = (hamiltonian, spin, freqs[-1])
            self._last_call = (hamiltonian, spin, freqs[-1])

    _last_call = None

    def _response_from_explicit_inverse(self):
        """rsp = G^-1 . rhs with a caller-supplied inverse (device GEMM)."""
        for op in self.operators:
            sva = op.supervector_dev("alph")
            if not self.is_uhf:
                Ginv = nv.to_dev(self._explicit_hessian_inv)
                self._append_rspvecs(op, "alph", nv.matmul(sva, Ginv, transpose_b=True))
            else:
                svb = op.supervector_dev("beta")
                _, G_ab, G_ba, _ = (nv.to_dev(g) for g in self._explicit_hessian)
                G_aa_inv, G_bb_inv = (nv.to_dev(g) for g in self._explicit_hessian_inv)
                torch = nv.torch_mod()
                t_b = nv.matmul(nv.matmul(svb, G_bb_inv, transpose_b=True), G_ab, transpose_b=True)
                t_a = nv.matmul(nv.matmul(sva, G_aa_inv, transpose_b=True), G_ba, transpose_b=True)
                right_a = eng.axpby(-1.0, t_b, 1.0, sva.clone())
                right_b = eng.axpby(-1.0, t_a, 1.0, svb.clone())
                self._append_rspvecs(op, "alph", nv.matmul(right_a, nv.to_dev(self.left_alph),
                                                          transpose_b=True))
                self._append_rspvecs(op, "beta", nv.matmul(right_b, nv.to_dev(self.left_beta),
                                                          transpose_b=True))
                del torch


class ExactInv(ExactLineqSolver):
    """Exact response vectors; semantics of ``np.linalg.inv(G) @ rhs`` for any frequency
    (reference solvers.py:483-520)."""

    _allow_lu_fallback = True

    def __init__(self, mocoeffs, moenergies, occupations, *,
                 inv_func: Callable[[np.ndarray], np.ndarray] | None = None) -> None:
        super().__init__(mocoeffs, moenergies, occupations)
        if inv_func is not None:
            self.inv_func = inv_func
        else:
            self.inv_func = np.linalg.inv

    def invert_explicit_hessian(self) -> None:
        assert self.explicit_hessian is not None
        if self.inv_func is np.linalg.inv:
            self._invert_with(self._gpu_inverse)
        else:
            self._invert_with(self.inv_func)


class ExactInvCholesky(ExactLineqSolver):
    """Cholesky-only variant: raises ``numpy.linalg.LinAlgError`` when G(w) is not positive
    definite, as ``scipy.linalg.cholesky`` does in the reference (solvers.py:523-555)."""

    _allow_lu_fallback = False

    def invert_explicit_hessian(self) -> None:
        assert self.explicit_hessian is not None
        self._invert_with(self._gpu_inverse_cholesky)


class IterativeLinEqSolver(LineqSolver):
    """Iterative CPHF (reference solvers.py:558-660), RHF only.

    Two methods:

    * ``method="fixed_point"`` (default when a ``jk_generator`` is given): the reference's loop,
      AO-direct with J/K builds.  With the GPU ``JKDense`` both J/K builds of every operator
      component are ONE pass of ``pmr_jk_contract`` over the AO ERI tensor per iteration (the
      reference makes 2 * ncomp separate JK calls, solvers.py:618-619).  Same initial guess, same
      update, same signed convergence measure, same iteration counts.
    * ``method="cg"`` (default when ``jk_generator`` is None): matrix-free block preconditioned
      conjugate gradients on ``[[K,-w],[-w,M]] [p;m] = rhs`` with sigma applied from the MO blocks
      (ia|jb), (ij|ab) by DMMA GEMMs -- no orbital Hessian, no AO tensor in the loop; every operator
      component and frequency is a column of ONE block iteration.  The MO blocks may be sharded over
      ranks by rows (``tei_mo_rows``): each rank applies its rows of sigma, the row blocks are
      all-gathered once per iteration.  Response vectors follow the reference iterative solver's
      convention (its X/Y halves correspond to the exact solver's at -w), so
      either method is a drop-in for the other.
    """

    def __init__(self, mocoeffs, moenergies, occupations, jk_generator: JK | None = None, *,
                 maxiter: int = 40, conv: float = 1.0e-9, method: str | None = None) -> None:
        super().__init__(mocoeffs, moenergies, occupations)
        self.jk_generator = jk_generator
        self.maxiter = maxiter
        self.conv = conv
        self.iterations = []
        if method is None:
            method = "fixed_point" if jk_generator is not None else "cg"
        assert method in ("fixed_point", "cg")
        self.method = method
        #: (i_lo, i_hi): ``tei_mo`` holds only these occupied rows of (ia|jb) and (ij|ab)
        self.tei_mo_rows = None
        self.residuals = []

    def run(self, hamiltonian: Hamiltonian, spin: Spin, program: Program | None,
            program_obj: Any) -> None:
        assert self.frequencies
        if self.method == "cg":
            if self.is_uhf:
                raise RuntimeError
            if not self.tei_mo:
                if program is None:
                    msg = "Program must be specified for computing 2-electron integrals in MO basis"
                    raise RuntimeError(msg)
                self.form_tei_mo(program, program_obj, AO2MOTransformationType.partial)
            self._run_cg(hamiltonian, spin, [float(w) for w in self.frequencies])
            return
        for frequency in self.frequencies:
            self._run(hamiltonian, spin, frequency, self.maxiter, self.conv)

    # ------------------------------------------------------------------ matrix-free block PCG
    def _run_cg(self, hamiltonian: Hamiltonian, spin: Spin, freqs) -> None:
        import ctypes as C

        from pymolresponse_b200 import distributed as pd

        assert self.tei_mo_type == AO2MOTransformationType.partial and len(self.tei_mo) == 2, (
            "the matrix-free path works from the partial blocks (ia|jb), (ij|ab)")
        torch = nv.torch_mod()
        L = nv.lib()
        nocc, nvirt, _, _ = self.occupations
        nov = nocc * nvirt
        ovov, oovv = (nv.to_dev(t) for t in self.tei_mo)
        i_lo, i_hi = self.tei_mo_rows if self.tei_mo_rows is not None else (0, nocc)
        ri = i_hi - i_lo
        assert tuple(ovov.shape) == (ri, nvirt, nocc, nvirt) and tuple(oovv.shape) == (ri, nocc, nvirt, nvirt)
        sharded = self.tei_mo_rows is not None and pd.rank_world()[1] > 1
        if sharded:
            assert (i_lo, i_hi) == pd.shard_range(nocc, *pd.rank_world()), (
                "row-sharded MO blocks must follow distributed.shard_range over the occupied index")
        else:
            assert (i_lo, i_hi) == (0, nocc), "a single rank needs all rows of the MO blocks"
        ops = list(self.operators)
        if not ops or nov == 0:
            return
        # columns: frequency-major, then operator, then component; reference-iterative convention:
        # its vectors at w are the exact solver's at -w (times -1 for imaginary operators)
        g_cols, signs, w_cols, flips = [], [], [], []
        for w in freqs:
            for op in ops:
                g = op.supervector_dev("alph")[:, :nov]
                s_ = +1.0 if op.is_imaginary else -1.0
                for c in range(int(g.shape[0])):
                    g_cols.append(g[c])
                    signs.append(s_)
                    w_cols.append(-w)
                    flips.append(-1.0 if op.is_imaginary else 1.0)
        k = len(g_cols)
        Gm = torch.stack(g_cols, dim=1).contiguous()                   # (nov, k)
        sg = torch.tensor(signs, dtype=torch.float64, device=Gm.device)
        wdev = torch.tensor(w_cols, dtype=torch.float64, device=Gm.device)
        b = torch.cat(((1.0 + sg)[None, :] * Gm, (1.0 - sg)[None, :] * Gm), dim=1).contiguous()
        eo, ev = bl.split_energies(self.moenergies[0], nocc)
        ediff = nv.empty(nov)
        nv.check(L.pmr_energy_differences(nv.ptr(eo), nocc, nv.ptr(ev), nvirt, nv.ptr(ediff),
                                          nv.stream_ptr()))
        c = {Spin.singlet: 2.0, Spin.triplet: 0.0}[spin]
        if hamiltonian == Hamiltonian.TDA:
            cK, sK, cM, sM = c, 0.0, c, 0.0
        else:
            cK, sK, cM, sM = 0.0, 1.0, 2.0 * c, -1.0
        need_U, need_W = (cK != 0.0 or cM != 0.0), (sK != 0.0 or sM != 0.0)
        k2 = 2 * k
        rows_loc = ri * nvirt
        U = nv.zeros(rows_loc, k2) if need_U else None
        V = nv.zeros(rows_loc, k2)
        W = nv.zeros(rows_loc, k2) if need_W else None
        Zp = nv.empty(nov, k2) if need_W else None
        q_loc = nv.empty(rows_loc, k2)
        null = C.c_void_p(0)

        def sigma(Z):
            """H Z for the (nov, 2k) block Z."""
            if rows_loc > 0:
                if need_U:   # (ia|jb) . Z : one GEMM, A k-contiguous
                    nv.dgemm(rows_loc, k2, nov, 1.0, ovov, 0, nov, 1, Z, 0, k2, 1, 0.0, U, 0, k2, 1)
                if need_W:   # sum_{b,j} (ib|ja) Z[j,b] : per i, A^T with K index (b,j), stride nvirt
                    nv.check(L.pmr_iter_permute_ov_rows(nv.ptr(Z), nocc, nvirt, k2, nv.ptr(Zp),
                                                        nv.stream_ptr()))
                    nv.dgemm(nvirt, k2, nov, 1.0, ovov, 0, 1, nvirt, Zp, 0, k2, 1, 0.0, W, 0, k2, 1,
                             batch1=ri, a_b=(nov * nvirt, 0), b_b=(0, 0), c_b=(nvirt * k2, 0))
                # sum_{j,b} (ij|ab) Z[j,b] : per (i, j) a (v x v) block, accumulated over j
                for j in range(nocc):
                    nv.dgemm(nvirt, k2, nvirt, 1.0, oovv, j * nvirt * nvirt, nvirt, 1,
                             Z, j * nvirt * k2, k2, 1, 0.0 if j == 0 else 1.0, V, 0, k2, 1,
                             batch1=ri, a_b=(nocc * nvirt * nvirt, 0), b_b=(0, 0), c_b=(nvirt * k2, 0))
                nv.check(L.pmr_iter_sigma_combine(
                    nv.ptr(U) if need_U else null, nv.ptr(V), nv.ptr(W) if need_W else null, nv.ptr(Z),
                    nv.ptr(ediff), nv.ptr(wdev), i_lo * nvirt, rows_loc, k, cK, sK, cM, sM,
                    nv.ptr(q_loc), nv.stream_ptr()))
            if sharded:
                return pd.allgather_rows(q_loc.view(ri, nvirt * k2), nocc).view(nov, k2)
            return q_loc

        work = nv.empty(int(L.pmr_iter_coldot_workspace(nov, k)))

        def dots(X, Y):
            out = nv.empty(k)
            nv.check(L.pmr_iter_coldot(nv.ptr(X), nv.ptr(Y), nov, k, nv.ptr(out), nv.ptr(work),
                                       nv.stream_ptr()))
            return out

        def precond(R):
            Y = nv.empty(nov, k2)
            nv.check(L.pmr_iter_precond(nv.ptr(R), nv.ptr(ediff), nv.ptr(wdev), nov, k, nv.ptr(Y),
                                        nv.stream_ptr()))
            return Y

        z = precond(b)                               # uncoupled guess (solvers.py:606-611)
        r = eng.axpby(-1.0, sigma(z), 1.0, b.clone())
        y = precond(r)
        d = y.clone()
        rho = dots(r, y)
        bnorm2 = dots(b, b).clamp_min(1e-300)
        converged, its, rel = False, 0, None
        # enough sweeps for machine precision below the first pole; ``maxiter`` keeps its meaning
        for it in range(1, max(self.maxiter, 1) + 1):
            its = it
            q = sigma(d)
            dq = dots(d, q)
            nv.check(L.pmr_iter_update_zr(nv.ptr(z), nv.ptr(r), nv.ptr(d), nv.ptr(q), nv.ptr(rho),
                                          nv.ptr(dq), nov, k, nv.stream_ptr()))
            rel = float(torch.sqrt((dots(r, r) / bnorm2).max()).item())   # one host read per sweep
            if rel * rel < self.conv:
                converged = True
                break
            y = precond(r)
            rho_new = dots(r, y)
            nv.check(L.pmr_iter_update_d(nv.ptr(d), nv.ptr(y), nv.ptr(rho_new), nv.ptr(rho), nov, k,
                                         nv.stream_ptr()))
            rho = rho_new
        self.iterations.append(its if converged else -1)
        self.residuals.append(rel)
        if not converged:
            return      # the reference silently appends nothing (solvers.py:638-660)
        flip = torch.tensor(flips, dtype=torch.float64, device=z.device)
        p_rows = (z[:, :k] * flip[None, :]).t().contiguous()      # (k, nov)
        m_rows = (z[:, k:] * flip[None, :]).t().contiguous()
        xy = nv.empty(k, 2 * nov)
        nv.check(L.pmr_pm_to_xy(nv.ptr(p_rows), nv.ptr(m_rows), k, nov, nov, nov, nv.ptr(xy), 2 * nov,
                                nv.stream_ptr()))
        col = 0
        for _w in freqs:
            for op in ops:
                nc = int(op.supervector_dev("alph").shape[0])
                self._append_rspvecs(op, "alph", xy[col:col + nc].contiguous())
                col += nc

    def _run(self, hamiltonian: Hamiltonian, spin: Spin, omega: float, maxiter: int,
             conv: float) -> None:
        nocc, nvirt, nocc_b, nvirt_b = self.occupations
        assert (nocc + nvirt) == (nocc_b + nvirt_b)
        norb = nocc + nvirt
        if self.is_uhf:
            raise RuntimeError
        torch = nv.torch_mod()
        L = nv.lib()
        Cd = nv.to_dev(self.mocoeffs[0])
        eps = nv.to_dev(self.moenergies[0]).diagonal().contiguous()
        Co = Cd[:, :nocc].contiguous()
        for operator in self.operators:
            O = nv.to_dev(operator.ao_integrals)
            ncomp = int(O.shape[0])
            # rhsmats[c] = C^T O_c C
            rhs = nv.empty(ncomp, norb, norb)
            for c in range(ncomp):
                rhs[c] = nv.matmul(nv.matmul(Cd, O[c].contiguous(), transpose_a=True), Cd)
            x = nv.zeros(ncomp, norb, norb)
            x_old = nv.zeros(ncomp, norb, norb)
            maxdiff = nv.empty(ncomp)
            # initial guess: uncoupled amplitudes rhs/(+-(e_a - e_i - w)) (solvers.py:606-611)
            nv.check(L.pmr_cphf_update(nv.ptr(x), nv.ptr(x_old), nv.ptr(rhs), nv.ptr(rhs),
                                       nv.ptr(eps), float(omega), norb, nocc, ncomp, 1,
                                       nv.ptr(maxdiff), nv.stream_ptr()))
            converged = False
            for it in range(1, maxiter + 1):
                G = nv.empty(ncomp, norb, norb)
                # densities D1 = Co (C x^T)[:, :o]^T, D2 = (-C x)[:, :o] Co^T, all components
                left, right = [], []
                for c in range(ncomp):
                    xc = x[c].contiguous()
                    left.append(Co)
                    right.append(nv.matmul(Cd, xc, transpose_b=True)[:, :nocc].contiguous())
                    cl2 = nv.matmul(Cd, xc)[:, :nocc].contiguous()
                    left.append(eng.axpby(-1.0, cl2, 0.0, nv.empty(cl2.shape[0], cl2.shape[1])))
                    right.append(Co)
                JK_list = self.jk_generator.compute_from_mocoeffs_batch(left, right)
                for c in range(ncomp):
                    (J1, K1), (J2, K2) = JK_list[2 * c], JK_list[2 * c + 1]
                    F = eng.axpby(2.0, J1, 0.0, nv.empty(J1.shape[0], J1.shape[1]))
                    eng.axpby(-1.0, K1, 1.0, F)
                    eng.axpby(2.0, J2, 1.0, F)
                    eng.axpby(-1.0, K2, 1.0, F)
                    G[c] = nv.matmul(nv.matmul(Cd, F, transpose_a=True), Cd)
                # x_old lags one sweep behind and starts at zero, as in the reference
                nv.check(L.pmr_cphf_update(nv.ptr(x), nv.ptr(x_old), nv.ptr(rhs), nv.ptr(G),
                                           nv.ptr(eps), float(omega), norb, nocc, ncomp, 0,
                                           nv.ptr(maxdiff), nv.stream_ptr()))
                rms = (maxdiff.cpu().numpy()) ** 2
                x_old.copy_(x)
                if rms.max() < conv:
                    self.iterations.append(it)
                    ov = x[:, :nocc, nocc:].reshape(ncomp, nocc * nvirt)                   # i*v + a
                    vo = x[:, nocc:, :nocc].transpose(1, 2).reshape(ncomp, nocc * nvirt)   # i*v + a
                    vec = torch.cat((ov, vo), dim=1).contiguous()
                    out = eng.axpby(-1.0, vec, 0.0, nv.empty(ncomp, 2 * nocc * nvirt))
                    self._append_rspvecs(operator, "alph", out)
                    converged = True
                    break
            if not converged:
                # the reference silently appends nothing (solvers.py:638-660)
                self.iterations.append(-1)


class EigSolver(Solver, ABC):
    """Base class for all solvers of the type AX = 0 (eigensolvers; reference solvers.py:663-690)."""

    _eigvals = None
    _eigvecs = None

    @property
    def eigvals(self):
        return self._eigvals

    @eigvals.setter
    def eigvals(self, x) -> None:
        self._eigvals = x

    @property
    def eigvecs(self):
        return self._eigvecs

    @eigvecs.setter
    def eigvecs(self, x) -> None:
        self._eigvecs = x

    @staticmethod
    def norm_xy(z, nocc: int, nvirt: int):
        """Scale z = [x; y] so that 2 (|x|^2 - |y|^2) = 1 (reference solvers.py:685-690)."""
        x, y = z.reshape(2, nvirt, nocc)
        norm = 2 * (np.linalg.norm(x) ** 2 - np.linalg.norm(y) ** 2)
        norm = 1 / np.sqrt(norm)
        return (x * norm).flatten(), (y * norm).flatten()


class EigSolverTDA(EigSolver, ABC):
    """Base class for eigensolvers within the Tamm-Dancoff approximation (solvers.py:693-696)."""


def _jacobi_svd(Gt, max_sweeps: int = 60, tol: float = 1e-15):
    """One-sided Jacobi SVD on the GPU (pmr_jacobi_svd).  ``Gt``: (N, N) CUDA tensor whose ROWS are
    the columns of G; overwritten.  Returns (sigma, gv, Vt, sweeps): singular values, g_p . v_p and
    the right singular vectors as rows of Vt, all unsorted device tensors."""
    import ctypes as C

    N = int(Gt.shape[0])
    Vt = nv.zeros(N, N)
    eng.shift_diagonal(Vt, N, 1.0)
    sigma, gv, scratch = nv.empty(max(1, N)), nv.empty(max(1, N)), nv.zeros(2)
    sweeps, off = C.c_int(0), C.c_double(0.0)
    nv.check(nv.lib().pmr_jacobi_svd(nv.ptr(Gt), nv.ptr(Vt), N, int(Gt.stride(0)) if N else 1,
                                     int(max_sweeps), float(tol), nv.ptr(sigma), nv.ptr(gv),
                                     nv.ptr(scratch), C.byref(sweeps), C.byref(off), nv.stream_ptr()))
    return sigma[:N], gv[:N], Vt, int(sweeps.value)


def _lower_factor(S):
    """Cholesky factor of an SPD device matrix as a plain lower-triangular GEMM operand."""
    Lf = S.clone()
    _dinv, info = eng.potrf(Lf)
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError(
            "A+B / A-B is not positive definite (unstable reference wavefunction): the symmetric "
            "reduction of the RPA eigenproblem does not apply")
    N = int(Lf.shape[0])
    nv.check(nv.lib().pmr_zero_upper(nv.ptr(Lf), N, int(Lf.stride(0)), nv.stream_ptr()))
    return Lf


class ExactDiagonalizationSolver(EigSolver, _HessianBlocks):
    """All excitation energies and eigenvectors of the RPA (or TDA) Hamiltonian
    [[A, B], [-B, -A]] (reference solvers.py:699-788, which calls scipy.linalg.eig on the 2N x 2N
    non-symmetric matrix).

    On the GPU the problem is reduced to a symmetric one of half the size: with K = A + B = L L^T
    and M = A - B = L_M L_M^T the squared excitation energies are the eigenvalues of L^T M L, i.e.
    the excitation energies are the SINGULAR VALUES of G = L_M^T L, computed by a one-sided Jacobi
    SVD kernel; X - Y = L v / w and X + Y = M (X - Y) / w follow by two GEMMs.  ``eigvals`` is real
    (the reference returns complex numbers with zero imaginary part), ascending; ``eigvecs`` has
    unit-norm columns [X; Y] like LAPACK's, signs fixed so that the largest component is positive.
    Needs a stable reference (A +- B positive definite); UHF raises as in the reference."""

    device_eig_results = False

    def form_explicit_hessian(self, hamiltonian: Hamiltonian, spin: Spin,
                              frequency: float | None) -> None:
        assert self.tei_mo is not None
        assert len(self.tei_mo) in (1, 2, 4, 6)
        assert isinstance(self.tei_mo_type, AO2MOTransformationType)
        if self.is_uhf:
            raise NotImplementedError  # as the reference (solvers.py:754-756)
        M, K, _, _ = self._assemble_rhf(hamiltonian, spin, False)
        self._MK = (M, K)
        self._lazy_hessian = None

    @property
    def explicit_hessian(self):
        """[[A, B], [-B, -A]] on the host, materialised on first access (solvers.py:751-752)."""
        if self._lazy_hessian is None and getattr(self, "_MK", None) is not None:
            M, K = (nv.to_host(t) for t in self._MK)
            A, B = 0.5 * (K + M), 0.5 * (K - M)
            self._lazy_hessian = np.block([[A, B], [-B, -A]])
        return self._lazy_hessian

    @explicit_hessian.setter
    def explicit_hessian(self, value) -> None:
        self._lazy_hessian = value

    def diagonalize_explicit_hessian(self) -> None:
        if self.is_uhf:
            raise NotImplementedError
        torch = nv.torch_mod()
        M, K = self._MK
        N = int(M.shape[0])
        Lk = _lower_factor(K)
        Lm = Lk if M is K else _lower_factor(M)
        # Gt = G^T = L^T L_M  (rows of Gt = columns of G = L_M^T L)
        Gt = nv.matmul(Lk, Lm, transpose_a=True)
        sigma, _gv, Vt, sweeps = _jacobi_svd(Gt)
        self.jacobi_sweeps = sweeps
        order = torch.argsort(sigma)
        w = sigma[order].contiguous()
        V = Vt[order].t().contiguous()                 # columns = eigenvectors of L^T M L, ascending
        D = nv.matmul(Lk, V)                           # X - Y (up to 1/w)
        S = nv.matmul(M, D)                            # w (X + Y) (same scale)
        winv = (1.0 / w).contiguous()
        nv.check(nv.lib().pmr_scale_columns(nv.ptr(S), N, N, N, nv.ptr(winv), nv.stream_ptr()))
        Z = nv.empty(2 * N, N)
        eng.axpby(0.5, S, 0.0, Z[:N])
        eng.axpby(0.5, D, 1.0, Z[:N])
        eng.axpby(0.5, S, 0.0, Z[N:])
        eng.axpby(-0.5, D, 1.0, Z[N:])
        nv.check(nv.lib().pmr_normalize_columns(nv.ptr(Z), 2 * N, N, N, nv.stream_ptr()))
        self._eigvals_dev, self._eigvecs_dev = w, Z
        self.eigvals = w if self.device_eig_results else nv.to_host(w)
        self.eigvecs = Z if self.device_eig_results else nv.to_host(Z)

    def run(self, hamiltonian: Hamiltonian, spin: Spin, program: Program | None,
            program_obj: Any) -> None:
        if not self.tei_mo:
            if program is None:
                msg = "Program must be specified for computing 2-electron integrals in MO basis"
                raise RuntimeError(msg)
            self.form_tei_mo(program, program_obj, AO2MOTransformationType.partial)
        self.form_explicit_hessian(hamiltonian, spin, None)
        self.diagonalize_explicit_hessian()


class ExactDiagonalizationSolverTDA(ExactDiagonalizationSolver, EigSolverTDA):
    """All eigenpairs of the TDA Hamiltonian A (reference solvers.py:791-856).  A is symmetric: its
    SVD by the same Jacobi kernel gives |eigenvalue| and the eigenvector, g_p . v_p the sign."""

    def form_explicit_hessian(self, hamiltonian: Hamiltonian, spin: Spin,
                              frequency: float | None) -> None:
        assert hamiltonian == Hamiltonian.TDA
        super().form_explicit_hessian(hamiltonian, spin, frequency)

    @property
    def explicit_hessian(self):
        if self._lazy_hessian is None and getattr(self, "_MK", None) is not None:
            self._lazy_hessian = nv.to_host(self._MK[0])
        return self._lazy_hessian

    @explicit_hessian.setter
    def explicit_hessian(self, value) -> None:
        self._lazy_hessian = value

    def diagonalize_explicit_hessian(self) -> None:
        if self.is_uhf:
            raise NotImplementedError
        torch = nv.torch_mod()
        A = self._MK[0]
        N = int(A.shape[0])
        Gt = A.t().contiguous()
        _sigma, gv, Vt, sweeps = _jacobi_svd(Gt)
        self.jacobi_sweeps = sweeps
        order = torch.argsort(gv)
        w = gv[order].contiguous()
        Z = Vt[order].t().contiguous()
        nv.check(nv.lib().pmr_normalize_columns(nv.ptr(Z), N, N, N, nv.stream_ptr()))
        self._eigvals_dev, self._eigvecs_dev = w, Z
        self.eigvals = w if self.device_eig_results else nv.to_host(w)
        self.eigvecs = Z if self.device_eig_results else nv.to_host(Z)
