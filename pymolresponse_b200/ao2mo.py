"""AO -> MO transformation of two-electron integrals on the GPU (FP64 DMMA GEMM chains).

Drop-in for the reference's ``pymolresponse/ao2mo.py`` ``AO2MO`` class: same constructor, same
``transform`` static method, same ``perform_*`` methods filling ``self.tei_mo`` with tuples of the
same length, order, shapes and C-order layout.  Differences, all additive:

* ``perform_uhf_partial`` / ``perform_uhf_full`` are implemented (the reference's base class raises
  NotImplementedError, ao2mo.py:72-78); tuple orders follow ``AO2MOpyscf``
  (interfaces/pyscf/ao2mo.py:69 and :102-109).
* the partial transforms use the occupied-first chain of ``pmr_ao2mo_partial`` (one shared first
  quarter per bra spin; 2.05x fewer flops than two independent chains, ao2mo.py:64-69);
* ``I`` may be a NumPy array or a CUDA tensor; with ``device_results=True`` the MO blocks stay in
  HBM as torch tensors (what the solvers consume directly).
* ``nu_range=(lo, hi)`` marks ``I`` as the slab ``I[:, lo:hi, :, :]`` of a tensor sharded over the
  second index of the first AO pair; with ``reduce=True`` the blocks are completed across the
  ``torch.distributed`` ranks (all-reduce of (ia|jb); reduce-scatter / finish / all-gather for
  (ij|ab)), otherwise they are this slab's partial sums.
* ``df_factor=B``: instead of the n^4 tensor the caller hands a factor
  ``B[P, mu, nu]`` with ``(mu nu|la si) = sum_P B[P,mu,nu] B[P,la,si]`` (density fitting, Cholesky
  vectors; the synthetic generator is one).  The MO blocks are then ``(ia|jb) = B_ia^T B_jb`` after
  two half-transforms of the factor: O(naux n^2 norb) flops and no n^4 object at all.
* ``ao_symmetry="s8"``: the caller states that a host-side ``I`` has the 8-fold permutational
  symmetry of real-orbital ERIs (what pyscf ``int2e`` and psi4 ``ao_eri`` return); only its unique
  part is copied to the GPU and the tensor is completed in HBM (a sixth of the bytes for the whole
  tensor, half for a ``nu_range`` slab).  The default ``"s1"`` copies the whole array and assumes
  nothing.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from pymolresponse_b200 import _native as nv
from pymolresponse_b200.utils import fix_mocoeffs_shape


#: stage times of the last partial transforms when ``_engine.STAGE_TIMES`` is on ({"ms": {...}})
LAST_STAGE: dict = {}


def _transform_dev(I, C1, C2, C3, C4):
    """I (n,n,n,n) and Ck (n, dk): contiguous CUDA tensors -> (d1,d2,d3,d4) CUDA tensor."""
    n = int(I.shape[0])
    d1, d2, d3, d4 = (int(c.shape[1]) for c in (C1, C2, C3, C4))
    L = nv.lib()
    out = nv.empty(d1, d2, d3, d4)
    if min(d1, d2, d3, d4) == 0:
        return out
    ws_n = int(L.pmr_ao2mo_transform_workspace(n, d1, d2, d3, d4))
    ws = nv.empty(ws_n)
    nv.check(L.pmr_ao2mo_transform(
        nv.ptr(I), n,
        nv.ptr(C1), int(C1.stride(0)), d1, nv.ptr(C2), int(C2.stride(0)), d2,
        nv.ptr(C3), int(C3.stride(0)), d3, nv.ptr(C4), int(C4.stride(0)), d4,
        nv.ptr(out), nv.ptr(ws), ws_n, nv.stream_ptr()))
    return out


def partial_blocks_dev(I_slab, nu_lo, Cbra, nocc_bra, kets, want_oovv, nu_block=0, reduce=False):
    """Occupied-first partial transform for one bra spin.

    I_slab: (n, nu_cnt, n, n) CUDA tensor; Cbra: (n, norb) CUDA tensor; kets: list of
    (Cket (n,norb), nocc_ket).  Returns ([ovov_k ...], oovv or None) as CUDA tensors.

    ``reduce=True`` (ranks hold different nu slabs): the (ia|jb) partial sums are all-reduced; the
    (ij|ab) chain stops after its two occupied quarters, the half-transformed (ij|la si) is
    reduce-scattered over i, every rank finishes the two virtual quarters for its own i only and the
    rows are all-gathered -- the o^2 n^2 (v + v^2/n) tail flops are not repeated on every rank.

    ``reduce="scatter"`` (row-block-sharded Hessian): every block comes back as this rank's occupied
    rows ``distributed.shard_range(o1, rank, world)`` only -- (ia|jb) is reduce-scattered over i
    (half the traffic of the all-reduce) and the all-gather of (ij|ab) is dropped.
    """
    from pymolresponse_b200 import _engine as eng
    from pymolresponse_b200 import distributed as pd

    clock = eng._StageClock(LAST_STAGE)
    clock.stage("ao2mo_transform")
    L = nv.lib()
    n, nu_cnt = int(I_slab.shape[0]), int(I_slab.shape[1])
    norb = int(Cbra.shape[1])
    o1, v1 = int(nocc_bra), norb - int(nocc_bra)
    ldc1 = int(Cbra.stride(0))
    kd = (nv.KetDesc * max(1, len(kets)))()
    outs = []
    keep = []
    for k, (Ck, ok) in enumerate(kets):
        o2, v2 = int(ok), int(Ck.shape[1]) - int(ok)
        out = nv.zeros(o1, v1, o2, v2)
        outs.append(out)
        keep.append(Ck)
        kd[k].Co2 = nv.ptr(Ck)
        kd[k].Cv2 = nv.ptr(Ck, o2)
        kd[k].ldc2 = int(Ck.stride(0))
        kd[k].o2, kd[k].v2 = o2, v2
        kd[k].ovov = nv.ptr(out)
    scatter = (reduce == "scatter") and pd.rank_world()[1] > 1
    split_tail = bool(reduce and want_oovv and pd.rank_world()[1] > 1)
    if split_tail:
        oovv = nv.zeros(o1, o1, n, n)       # receives the half-transformed (i j|la si)
    else:
        oovv = nv.zeros(o1, o1, v1, v1) if want_oovv else None
    ws_n = int(L.pmr_ao2mo_partial_workspace(n, nu_cnt, o1, v1, len(kets), kd, int(want_oovv),
                                             int(nu_block)))
    ws = nv.empty(ws_n)
    nv.check(L.pmr_ao2mo_partial(
        nv.ptr(I_slab), n, int(nu_lo), nu_cnt, nv.ptr(Cbra), nv.ptr(Cbra, o1), ldc1, o1, v1,
        len(kets), kd, nv.ptr(oovv) if want_oovv else C.c_void_p(0), 2 if split_tail else 0,
        int(nu_block), nv.ptr(ws), ws_n, nv.stream_ptr()))
    del ws
    clock.mark("transform")
    clock.stage("ao2mo_collectives")
    if reduce:
        if scatter:
            outs = [pd.reduce_scatter_rows(t, o1) for t in outs]
            clock.mark("reduce_scatter_ovov")
        else:
            pd.allreduce_blocks(outs)
            clock.mark("allreduce_ovov")
        if split_tail:
            Y_loc = pd.reduce_scatter_rows(oovv, o1)            # (rows, o1, n, n), summed over ranks
            clock.mark("reduce_scatter_oovv")
            del oovv
            rows = int(Y_loc.shape[0])
            loc = nv.zeros(rows, o1, v1, v1)
            if rows > 0 and v1 > 0:
                tw = int(L.pmr_ao2mo_oovv_tail_workspace(n, rows, o1, v1))
                tws = nv.empty(tw)
                nv.check(L.pmr_ao2mo_oovv_tail(nv.ptr(Y_loc), n, nv.ptr(Cbra, o1), ldc1, rows, o1, v1,
                                               nv.ptr(loc), 0, nv.ptr(tws), tw, nv.stream_ptr()))
            clock.mark("oovv_tail")
            if scatter:
                oovv = loc
            else:
                oovv = pd.allgather_rows(loc, o1)
                clock.mark("allgather_oovv")
        elif want_oovv:
            pd.allreduce_blocks([oovv])
    clock.finish()
    return outs, oovv


def factor_to_mo(B, Cs):
    """B (naux, n, n), Cs (n, norb) CUDA tensors -> Bmo (naux, norb, norb) with
    Bmo[P, p, q] = sum_{mu nu} Cs[mu, p] B[P, mu, nu] Cs[nu, q] (two GEMM stages)."""
    naux, n = int(B.shape[0]), int(B.shape[1])
    norb = int(Cs.shape[1])
    ldc = int(Cs.stride(0))
    T = nv.empty(naux, n, norb)
    nv.dgemm(naux * n, norb, n, 1.0, B, 0, n, 1, Cs, 0, ldc, 1, 0.0, T, 0, norb, 1)
    Bmo = nv.empty(naux, norb, norb)
    # per P: Bmo_P = Cs^T T_P
    nv.dgemm(norb, norb, n, 1.0, Cs, 0, 1, ldc, T, 0, norb, 1, 0.0, Bmo, 0, norb, 1,
             batch1=naux, a_b=(0, 0), b_b=(n * norb, 0), c_b=(norb * norb, 0))
    return Bmo


def _factor_block(Bmo, p0, p1, q0, q1):
    """Contiguous (naux, (p1-p0)*(q1-q0)) copy of the sub-block Bmo[:, p0:p1, q0:q1]."""
    naux = int(Bmo.shape[0])
    return Bmo[:, p0:p1, q0:q1].contiguous().view(naux, (p1 - p0) * (q1 - q0))


def _factor_product(X, Y, shape):
    """out[x, y] = sum_P X[P, x] Y[P, y] -> tensor of ``shape`` (both operands mn-contiguous: the
    TMA-bulk GEMM kernel)."""
    naux, dx, dy = int(X.shape[0]), int(X.shape[1]), int(Y.shape[1])
    out = nv.empty(*shape)
    if dx and dy:
        nv.dgemm(dx, dy, naux, 1.0, X, 0, 1, dx, Y, 0, dy, 1, 0.0, out, 0, dy, 1)
    return out


class AO2MO:
    """Interface for AO-to-MO transformations of two-electron integrals (reference ao2mo.py:15-78)."""

    def __init__(self, C, occupations, I=None, *, device_results: bool = False,  # noqa: E741
                 nu_range: tuple[int, int] | None = None, nu_block: int = 0,
                 reduce: bool = False, ao_symmetry: str = "s1", df_factor=None) -> None:
        self.C = fix_mocoeffs_shape(C) if not nv.is_device_array(C) else (C if C.dim() == 3 else C[None])
        self.occupations = occupations
        self.I = I
        self.nocc_alph, self.nvirt_alph, self.nocc_beta, self.nvirt_beta = occupations
        self.tei_mo = tuple()
        self.device_results = device_results
        self.nu_range = nu_range
        self.nu_block = nu_block
        # complete the blocks over torch.distributed ranks (nu-sharded I): True = every rank gets the
        # whole blocks, "scatter" = every rank keeps its occupied rows only (see ``row_ranges``)
        assert reduce in (False, True, "scatter")
        self.reduce = reduce
        #: per spin, the occupied rows [lo, hi) the blocks of ``tei_mo`` hold (None: all rows)
        self.row_ranges = None
        assert ao_symmetry in ("s1", "s8")
        self.ao_symmetry = ao_symmetry
        self.df_factor = df_factor
        if df_factor is not None:
            assert I is None and nu_range is None, "give either the AO tensor or its factor"

    # ------------------------------------------------------------------ helpers
    def _out(self, tensors):
        if self.device_results:
            return tuple(tensors)
        return tuple(nv.to_host(t) for t in tensors)

    def _I_dev(self):
        assert self.I is not None
        if self.ao_symmetry == "s8" and not nv.is_device_array(self.I):
            if self.nu_range is None:
                return nv.upload_eri_s8(self.I)
            return nv.upload_eri_slab_s2(self.I)   # a nu-slab: only (la si| = (si la| applies
        I = nv.to_dev(self.I)
        assert I.dim() == 4
        return I

    def _slab(self):
        I = self._I_dev()
        n = int(I.shape[0])
        if self.nu_range is None:
            assert I.shape == (n, n, n, n)
            return I, 0
        lo, hi = self.nu_range
        assert I.shape == (n, hi - lo, n, n)
        return I, lo

    def _C_dev(self, spin: int):
        return nv.to_dev(self.C[spin])

    def _B_dev(self):
        B = nv.to_dev(self.df_factor)
        assert B.dim() == 3 and B.shape[1] == B.shape[2]
        return B

    def _factor_blocks(self, spins):
        """Per spin: (B_ov (naux, o v), B_oo (naux, o o), B_vv (naux, v v), o, v)."""
        out = []
        for spin in spins:
            Cs = self._C_dev(spin)
            o = self.nocc_alph if spin == 0 else self.nocc_beta
            norb = int(Cs.shape[1])
            Bmo = factor_to_mo(self._B_dev(), Cs)
            out.append((_factor_block(Bmo, 0, o, o, norb), _factor_block(Bmo, 0, o, 0, o),
                        _factor_block(Bmo, o, norb, o, norb), o, norb - o))
        return out

    # ------------------------------------------------------------------ reference API
    @staticmethod
    def transform(I, C1, C2, C3, C4):  # noqa: E741
        """(AB|CD) = sum C1[mu,A] C2[nu,B] C3[la,C] C4[si,D] (mu nu|la si), quarters in index order
        1,2,3,4 (reference ao2mo.py:35-50)."""
        on_dev = nv.is_device_array(I)
        out = _transform_dev(nv.to_dev(I), nv.to_dev(C1), nv.to_dev(C2), nv.to_dev(C3), nv.to_dev(C4))
        return out if on_dev else nv.to_host(out)

    def perform_rhf_full(self) -> None:
        r"""(mu nu|la si) -> ((pq|rs),)  (reference ao2mo.py:52-56)."""
        assert self.nu_range is None, "full transforms need the whole AO tensor"
        Ca = self._C_dev(0)
        if self.df_factor is not None:
            norb = int(Ca.shape[1])
            Bf = _factor_block(factor_to_mo(self._B_dev(), Ca), 0, norb, 0, norb)
            self.tei_mo = self._out((_factor_product(Bf, Bf, (norb,) * 4),))
            return
        I = self._I_dev()
        self.tei_mo = self._out((_transform_dev(I, Ca, Ca, Ca, Ca),))

    def perform_rhf_partial(self) -> None:
        r"""(mu nu|la si) -> ((ia|jb), (ij|ab))  (reference ao2mo.py:58-70)."""
        if self.df_factor is not None:
            (ov, oo, vv, o, v), = self._factor_blocks((0,))
            self.tei_mo = self._out((_factor_product(ov, ov, (o, v, o, v)),
                                     _factor_product(oo, vv, (o, o, v, v))))
            return
        I, lo = self._slab()
        Ca = self._C_dev(0)
        (ovov,), oovv = partial_blocks_dev(I, lo, Ca, self.nocc_alph, [(Ca, self.nocc_alph)], True,
                                           self.nu_block, self.reduce)
        self._set_row_ranges()
        self.tei_mo = self._out((ovov, oovv))

    def perform_uhf_full(self) -> None:
        r"""-> (aaaa, aabb, bbaa, bbbb), each (n,n,n,n)  (interfaces/pyscf/ao2mo.py:49-69)."""
        assert self.nu_range is None, "full transforms need the whole AO tensor"
        Ca, Cb = self._C_dev(0), self._C_dev(1)
        if self.df_factor is not None:
            na_, nb_ = int(Ca.shape[1]), int(Cb.shape[1])
            Fa = _factor_block(factor_to_mo(self._B_dev(), Ca), 0, na_, 0, na_)
            Fb = _factor_block(factor_to_mo(self._B_dev(), Cb), 0, nb_, 0, nb_)
            self.tei_mo = self._out((
                _factor_product(Fa, Fa, (na_,) * 4), _factor_product(Fa, Fb, (na_, na_, nb_, nb_)),
                _factor_product(Fb, Fa, (nb_, nb_, na_, na_)), _factor_product(Fb, Fb, (nb_,) * 4)))
            return
        I = self._I_dev()
        self.tei_mo = self._out((
            _transform_dev(I, Ca, Ca, Ca, Ca),
            _transform_dev(I, Ca, Ca, Cb, Cb),
            _transform_dev(I, Cb, Cb, Ca, Ca),
            _transform_dev(I, Cb, Cb, Cb, Cb),
        ))

    def perform_uhf_partial(self) -> None:
        r"""-> (ovov_aaaa, ovov_aabb, ovov_bbaa, ovov_bbbb, oovv_aaaa, oovv_bbbb)
        (interfaces/pyscf/ao2mo.py:71-109)."""
        if self.df_factor is not None:
            (ova, ooa, vva, oa, va), (ovb, oob, vvb, ob, vb) = self._factor_blocks((0, 1))
            self.tei_mo = self._out((
                _factor_product(ova, ova, (oa, va, oa, va)), _factor_product(ova, ovb, (oa, va, ob, vb)),
                _factor_product(ovb, ova, (ob, vb, oa, va)), _factor_product(ovb, ovb, (ob, vb, ob, vb)),
                _factor_product(ooa, vva, (oa, oa, va, va)), _factor_product(oob, vvb, (ob, ob, vb, vb))))
            return
        I, lo = self._slab()
        Ca, Cb = self._C_dev(0), self._C_dev(1)
        na, nb = self.nocc_alph, self.nocc_beta
        (aaaa, aabb), oovv_a = partial_blocks_dev(I, lo, Ca, na, [(Ca, na), (Cb, nb)], True,
                                                  self.nu_block, self.reduce)
        (bbaa, bbbb), oovv_b = partial_blocks_dev(I, lo, Cb, nb, [(Ca, na), (Cb, nb)], True,
                                                  self.nu_block, self.reduce)
        self._set_row_ranges()
        self.tei_mo = self._out((aaaa, aabb, bbaa, bbbb, oovv_a, oovv_b))

    def _set_row_ranges(self):
        from pymolresponse_b200 import distributed as pd

        rank, world = pd.rank_world()
        if self.reduce == "scatter" and world > 1:
            self.row_ranges = (pd.shard_range(self.nocc_alph, rank, world),
                               pd.shard_range(self.nocc_beta, rank, world))
        else:
            self.row_ranges = None


def flops_rhf_partial_df(n: int, o: int, naux: int) -> float:
    """Algorithmic flops of the factor route: two half-transforms of B, then the two products."""
    v = n - o
    return 2.0 * naux * n * n * n + 2.0 * naux * n * n * n + 2.0 * naux * (o * v) ** 2 + 2.0 * naux * o * o * v * v


def flops_rhf_partial(n: int, o: int) -> float:
    """Algorithmic FP64 flop count of the occupied-first RHF partial transform."""
    v = n - o
    return (2.0 * n**4 * o + 2 * 2.0 * n**3 * o**2 + 2 * 2.0 * n**2 * o**2 * v
            + 2 * 2.0 * n * o**2 * v**2)


_ = np  # numpy is part of the public contract of this module (arrays in, arrays out)
