r"""Orbital-Hessian blocks from partially transformed MO integrals, :math:`(ia|jb)` and
:math:`(ij|ab)`, assembled on the GPU.

Drop-in for the reference's ``pymolresponse/explicit_equations_partial.py`` (same eight function
names, argument order and return shapes).  NumPy arrays in -> NumPy array out; CUDA tensors in ->
CUDA tensor out.  All eight functions are one launch of the gather kernel ``pmr_assemble``;
the values are bit-identical to the reference's NumPy arithmetic (same operation order, no FMA
contraction).  Compound index: ``ia = i*nvirt + a`` (virtual fast).
"""

from __future__ import annotations

from pymolresponse_b200 import _blocks as bl
from pymolresponse_b200 import _native as nv


def _ret(t, on_dev):
    return t if on_dev else nv.to_host(t)


def _check_ovov_oovv(ovov_shape, oovv_shape):
    assert len(ovov_shape) == len(oovv_shape) == 4
    assert ovov_shape[0] == ovov_shape[2] == oovv_shape[0] == oovv_shape[1]
    assert ovov_shape[1] == ovov_shape[3] == oovv_shape[2] == oovv_shape[3]


def _check_E(E_MO, norb):
    assert len(E_MO.shape) == 2
    assert E_MO.shape[0] == E_MO.shape[1] == norb


def _a_partial(kind, E_MO, ovov, oovv):
    on_dev = nv.is_device_array(oovv)
    oovv_d = nv.to_dev(oovv)
    ovov_d = nv.to_dev(ovov) if ovov is not None else None
    nocc, nvirt = int(oovv_d.shape[0]), int(oovv_d.shape[2])
    _check_E(E_MO, nocc + nvirt)
    eps = bl.split_energies(E_MO, nocc)
    A, _ = bl.run(nocc, nvirt, nocc, nvirt, bl.partial_terms(kind, ovov_d, oovv_d), eps, "P")
    return _ret(A, on_dev)


def _b_partial(kind, ovov):
    on_dev = nv.is_device_array(ovov)
    ovov_d = nv.to_dev(ovov)
    shape = ovov_d.shape
    assert len(shape) == 4
    assert shape[0] == shape[2]
    assert shape[1] == shape[3]
    nocc, nvirt = int(shape[0]), int(shape[1])
    _, B = bl.run(nocc, nvirt, nocc, nvirt, bl.partial_terms(kind, ovov_d, None), None, "Q")
    return _ret(B, on_dev)


def form_rpa_a_matrix_mo_singlet_partial(E_MO, TEI_MO_iajb, TEI_MO_ijab):
    r"""A (CIS) matrix, singlet: :math:`2(ia|jb) - (ij|ab) + \delta(\epsilon_a-\epsilon_i)`
    (reference explicit_equations_partial.py:8-40)."""
    _check_ovov_oovv(TEI_MO_iajb.shape, TEI_MO_ijab.shape)
    return _a_partial("singlet", E_MO, TEI_MO_iajb, TEI_MO_ijab)


def form_rpa_a_matrix_mo_triplet_partial(E_MO, TEI_MO_ijab):
    r"""A matrix, triplet: :math:`-(ij|ab) + \delta(\epsilon_a-\epsilon_i)`
    (reference explicit_equations_partial.py:43-72)."""
    shape = TEI_MO_ijab.shape
    assert len(shape) == 4
    assert shape[0] == shape[1]
    assert shape[2] == shape[3]
    return _a_partial("triplet", E_MO, None, TEI_MO_ijab)


def form_rpa_b_matrix_mo_singlet_partial(TEI_MO_iajb):
    r"""B matrix (reference sign), singlet: :math:`-(2(ia|jb) - (ib|ja))`
    (reference explicit_equations_partial.py:75-96)."""
    return _b_partial("singlet", TEI_MO_iajb)


def form_rpa_b_matrix_mo_triplet_partial(TEI_MO_iajb):
    r"""B matrix (reference sign), triplet: :math:`+(ib|ja)`
    (reference explicit_equations_partial.py:99-118)."""
    return _b_partial("triplet", TEI_MO_iajb)


def form_rpa_a_matrix_mo_singlet_ss_partial(E_MO, TEI_MO_iajb, TEI_MO_ijab):
    r"""Same-spin part of A: :math:`(ia|jb) - (ij|ab) + \delta(\epsilon_a-\epsilon_i)`
    (reference explicit_equations_partial.py:121-151)."""
    _check_ovov_oovv(TEI_MO_iajb.shape, TEI_MO_ijab.shape)
    return _a_partial("ss", E_MO, TEI_MO_iajb, TEI_MO_ijab)


def form_rpa_a_matrix_mo_singlet_os_partial(TEI_MO_iajb_xxyy):
    r"""Opposite-spin part of A: :math:`(ia|jb)_{xxyy}` as a (nov_x, nov_y) matrix
    (reference explicit_equations_partial.py:154-171)."""
    on_dev = nv.is_device_array(TEI_MO_iajb_xxyy)
    t = nv.to_dev(TEI_MO_iajb_xxyy)
    assert t.dim() == 4
    ox, vx, oy, vy = (int(s) for s in t.shape)
    A, _ = bl.run(ox, vx, oy, vy, bl.partial_terms("os", t, None), None, "P")
    return _ret(A, on_dev)


def form_rpa_b_matrix_mo_singlet_ss_partial(TEI_MO_iajb):
    r"""Same-spin part of B (reference sign): :math:`-((ia|jb) - (ib|ja))`
    (reference explicit_equations_partial.py:174-194)."""
    return _b_partial("ss", TEI_MO_iajb)


def form_rpa_b_matrix_mo_singlet_os_partial(TEI_MO_iajb_xxyy):
    r"""Opposite-spin part of B (reference sign): :math:`-(ia|jb)_{xxyy}`
    (reference explicit_equations_partial.py:197-214)."""
    on_dev = nv.is_device_array(TEI_MO_iajb_xxyy)
    t = nv.to_dev(TEI_MO_iajb_xxyy)
    assert t.dim() == 4
    ox, vx, oy, vy = (int(s) for s in t.shape)
    _, B = bl.run(ox, vx, oy, vy, bl.partial_terms("os", t, None), None, "Q")
    return _ret(B, on_dev)
