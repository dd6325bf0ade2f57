"""Gather-term tables for the orbital-Hessian assembly kernel (pmr_assemble).

Every A/B block of the reference (explicit_equations_partial.py, explicit_equations_full.py) is
``P = c1*T1 - c2*T2 (+ diagonal)`` or ``Q = -(c1*U1 - c2*U2)`` with T/U being an MO integral
tensor read through a fixed index permutation.  This module writes those permutations down as
four element strides per operand; the CUDA kernel does the rest.
"""

from __future__ import annotations

from pymolresponse_b200 import _native as nv

ABSENT = None


# ``i_lo``: the tensor holds only the occupied rows [i_lo, i_lo + t.shape[0]) of the block (row-block
# sharded Hessian); the kernel addresses with the GLOBAL i, so the base pointer is moved back by
# i_lo rows (it is only dereferenced for i inside the range the launch covers).


def _ovov_direct(t, coef, i_lo=0):
    """(i,a,j,b) -> ovov[i,a,j,b]"""
    _, v, o2, v2 = t.shape
    s_i = v * o2 * v2
    return nv.term(t, -i_lo * s_i, s_i=s_i, s_a=o2 * v2, s_j=v2, s_b=1, coef=coef)


def _ovov_swap13(t, coef, i_lo=0):
    """(i,a,j,b) -> ovov[i,b,j,a]  (swapaxes(1,3), explicit_equations_partial.py:93)"""
    _, v, o, v2 = t.shape
    assert v == v2
    s_i = v * o * v
    return nv.term(t, -i_lo * s_i, s_i=s_i, s_b=o * v, s_j=v, s_a=1, coef=coef)


def _oovv_swap12(t, coef, i_lo=0):
    """(i,a,j,b) -> oovv[i,j,a,b]  (swapaxes(1,2), explicit_equations_partial.py:35)"""
    _, o, v, v2 = t.shape
    assert v == v2
    s_i = o * v * v
    return nv.term(t, -i_lo * s_i, s_i=s_i, s_j=v * v, s_a=v, s_b=1, coef=coef)


def _full(t, pattern, ox, oy, coef):
    """(i,a,j,b) -> TEI[...] for the four access patterns of explicit_equations_full.py."""
    n = t.shape[0]
    n2, n3 = n * n, n * n * n
    if pattern == "aijb":    # TEI[a+ox, i, j, b+oy]   (:34, :153, :182)
        return nv.term(t, ox * n3 + oy, s_a=n3, s_i=n2, s_j=n, s_b=1, coef=coef)
    if pattern == "abji":    # TEI[a+o, b+o, j, i]     (:35, :69)
        return nv.term(t, ox * n3 + oy * n2, s_a=n3, s_b=n2, s_j=n, s_i=1, coef=coef)
    if pattern == "aibj":    # TEI[a+ox, i, b+oy, j]   (:97, :207, :236)
        return nv.term(t, ox * n3 + oy * n, s_a=n3, s_i=n2, s_b=n, s_j=1, coef=coef)
    if pattern == "ajbi":    # TEI[a+o, j, b+o, i]     (:98, :122)
        return nv.term(t, ox * n3 + oy * n, s_a=n3, s_j=n2, s_b=n, s_i=1, coef=coef)
    raise ValueError(pattern)


_NONE = None


def absent():
    return nv.term(None)


def partial_terms(kind, ovov, oovv, i_lo=0):
    """(p1, p2, q1, q2) for kind in {'singlet','triplet','ss','os'}; tensors may be None when the
    corresponding output is not requested.  ``i_lo``: first occupied row held by row-sharded tensors."""
    a = absent()
    if kind == "singlet":
        c = 2.0
    elif kind in ("ss", "os"):
        c = 1.0
    elif kind == "triplet":
        c = 0.0
    else:
        raise ValueError(kind)
    if kind == "os":
        p1 = _ovov_direct(ovov, 1.0, i_lo) if ovov is not None else a
        return p1, absent(), (_ovov_direct(ovov, 1.0, i_lo) if ovov is not None else a), absent()
    p1 = _ovov_direct(ovov, c, i_lo) if (ovov is not None and c != 0.0) else a
    p2 = _oovv_swap12(oovv, 1.0, i_lo) if oovv is not None else absent()
    q1 = _ovov_direct(ovov, c, i_lo) if (ovov is not None and c != 0.0) else absent()
    q2 = _ovov_swap13(ovov, 1.0, i_lo) if ovov is not None else absent()
    return p1, p2, q1, q2


def full_terms(kind, tei, ox, oy):
    if kind == "singlet":
        c = 2.0
    elif kind in ("ss", "os"):
        c = 1.0
    elif kind == "triplet":
        c = 0.0
    else:
        raise ValueError(kind)
    if kind == "os":
        return (_full(tei, "aijb", ox, oy, 1.0), absent(), _full(tei, "aibj", ox, oy, 1.0), absent())
    p1 = _full(tei, "aijb", ox, oy, c) if c != 0.0 else absent()
    p2 = _full(tei, "abji", ox, oy, 1.0)
    q1 = _full(tei, "aibj", ox, oy, c) if c != 0.0 else absent()
    q2 = _full(tei, "ajbi", ox, oy, 1.0)
    return p1, p2, q1, q2


def split_energies(E_MO, nocc):
    """Diagonal of a (norb, norb) MO-energy matrix -> device (eps_occ, eps_virt)."""
    E = nv.to_dev(E_MO)
    assert E.dim() == 2 and E.shape[0] == E.shape[1]
    eps = E.diagonal().contiguous()
    return eps[:nocc].contiguous(), eps[nocc:].contiguous()


def run(no_x, nv_x, no_y, nv_y, terms, eps, want, out_mode=0, out0=None, out1=None, ld=None,
        off0=0, off1=0, i_lo=0, i_hi=None):
    """Launch the assembly kernel.  ``want`` in {'P','Q','PQ'} selects which outputs exist when
    out0/out1 are not supplied.  Returns (out0, out1) device tensors (None where not produced).
    ``[i_lo, i_hi)``: only these occupied rows are produced; supplied outputs then hold the row block
    alone ((i_hi - i_lo) * nv_x rows), as do the row-sharded input tensors behind ``terms``."""
    p1, p2, q1, q2 = terms
    novy = no_y * nv_y
    i_hi = no_x if i_hi is None else i_hi
    if out0 is None and out1 is None:
        if "P" in want:
            out0 = nv.empty((i_hi - i_lo) * nv_x, novy)
        if "Q" in want:
            out1 = nv.empty((i_hi - i_lo) * nv_x, novy)
        ld = novy
    if "P" not in want and out_mode == 0:
        p1, p2 = absent(), absent()
    if "Q" not in want and out_mode == 0:
        q1, q2 = absent(), absent()
    eo, ev = eps if eps is not None else (None, None)
    if (i_hi - i_lo) * nv_x * novy > 0:
        shift = -i_lo * nv_x * (ld if ld is not None else novy)   # outputs are addressed with global i
        nv.assemble(no_x, nv_x, no_y, nv_y, p1, p2, q1, q2, eo, ev, out_mode, out0, out1,
                    ld0=ld, ld1=ld, out0_off=off0 + shift, out1_off=off1 + shift, i_lo=i_lo, i_hi=i_hi)
    return out0, out1
