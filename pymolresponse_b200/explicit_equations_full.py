r"""Orbital-Hessian blocks from fully transformed MO integrals :math:`(pq|rs)`, assembled on the
GPU.

Drop-in for the reference's ``pymolresponse/explicit_equations_full.py`` (same eight function
names, argument order and return shapes).  The reference runs four nested Python loops per
block; here each block is one launch of the gather kernel ``pmr_assemble`` with the index
permutation expressed as element strides into the (n,n,n,n) tensor.  Values are bit-identical
to the reference's.  Compound index: ``ia = i*nvirt + a``.
"""

from __future__ import annotations

from pymolresponse_b200 import _blocks as bl
from pymolresponse_b200 import _native as nv


def _ret(t, on_dev):
    return t if on_dev else nv.to_host(t)


def _a_full(kind, E_MO, TEI_MO, nocc):
    on_dev = nv.is_device_array(TEI_MO)
    t = nv.to_dev(TEI_MO)
    norb = int(E_MO.shape[0])
    assert t.dim() == 4 and int(t.shape[0]) == norb
    nvirt = norb - nocc
    eps = bl.split_energies(E_MO, nocc)
    A, _ = bl.run(nocc, nvirt, nocc, nvirt, bl.full_terms(kind, t, nocc, nocc), eps, "P")
    return _ret(A, on_dev)


def _b_full(kind, TEI_MO, nocc):
    on_dev = nv.is_device_array(TEI_MO)
    t = nv.to_dev(TEI_MO)
    assert t.dim() == 4
    nvirt = int(t.shape[0]) - nocc
    _, B = bl.run(nocc, nvirt, nocc, nvirt, bl.full_terms(kind, t, nocc, nocc), None, "Q")
    return _ret(B, on_dev)


def form_rpa_a_matrix_mo_singlet_full(E_MO, TEI_MO, nocc):
    r"""A[ia,jb] = 2 TEI[a,i,j,b] - TEI[a,b,j,i] + diagonal (virtual indices offset by nocc;
    reference explicit_equations_full.py:8-40)."""
    return _a_full("singlet", E_MO, TEI_MO, nocc)


def form_rpa_a_matrix_mo_triplet_full(E_MO, TEI_MO, nocc):
    r"""A[ia,jb] = -TEI[a,b,j,i] + diagonal (reference explicit_equations_full.py:43-72)."""
    return _a_full("triplet", E_MO, TEI_MO, nocc)


def form_rpa_b_matrix_mo_singlet_full(TEI_MO, nocc):
    r"""B[ia,jb] = -(2 TEI[a,i,b,j] - TEI[a,j,b,i]) (reference explicit_equations_full.py:75-100)."""
    return _b_full("singlet", TEI_MO, nocc)


def form_rpa_b_matrix_mo_triplet_full(TEI_MO, nocc):
    r"""B[ia,jb] = +TEI[a,j,b,i] (reference explicit_equations_full.py:103-124)."""
    return _b_full("triplet", TEI_MO, nocc)


def form_rpa_a_matrix_mo_singlet_ss_full(E_MO, TEI_MO, nocc):
    r"""A_ss[ia,jb] = TEI[a,i,j,b] - TEI[a,b,j,i] + diagonal
    (reference explicit_equations_full.py:127-155)."""
    return _a_full("ss", E_MO, TEI_MO, nocc)


def form_rpa_a_matrix_mo_singlet_os_full(TEI_MO_xxyy, nocc_x, nocc_y):
    r"""A_os[ia,jb] = TEI_xxyy[a,i,j,b] (reference explicit_equations_full.py:158-184)."""
    on_dev = nv.is_device_array(TEI_MO_xxyy)
    t = nv.to_dev(TEI_MO_xxyy)
    vx, vy = int(t.shape[0]) - nocc_x, int(t.shape[2]) - nocc_y
    A, _ = bl.run(nocc_x, vx, nocc_y, vy, bl.full_terms("os", t, nocc_x, nocc_y), None, "P")
    return _ret(A, on_dev)


def form_rpa_b_matrix_mo_singlet_ss_full(TEI_MO, nocc):
    r"""B_ss[ia,jb] = -(TEI[a,i,b,j] - TEI[a,j,b,i])
    (reference explicit_equations_full.py:187-209)."""
    return _b_full("ss", TEI_MO, nocc)


def form_rpa_b_matrix_mo_singlet_os_full(TEI_MO_xxyy, nocc_x, nocc_y):
    r"""B_os[ia,jb] = -TEI_xxyy[a,i,b,j] (reference explicit_equations_full.py:212-238)."""
    on_dev = nv.is_device_array(TEI_MO_xxyy)
    t = nv.to_dev(TEI_MO_xxyy)
    vx, vy = int(t.shape[0]) - nocc_x, int(t.shape[2]) - nocc_y
    _, B = bl.run(nocc_x, vx, nocc_y, vy, bl.full_terms("os", t, nocc_x, nocc_y), None, "Q")
    return _ret(B, on_dev)
