"""CPU oracle: a NumPy restatement of pymolresponse's linear-response hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product package ``pymolresponse_b200`` never does;
its arithmetic runs in the CUDA library and fails loudly when that is missing.

Every function follows the reference's *algorithm* (same quarter order, same explicit
super-Hessian, same explicit inverse per frequency) so that timing it is a fair stand-in
for the reference's CPU path, and cites the reference file:line it restates.
Parity is PINNED: ``tests/test_oracle_golden.py`` checks every function here against
(1) the literals of the reference's own known-answer tests (water/STO-3G,
``tests/test_final_result.py:54-63,135-146,216-225,295-306``), (2) the DALTON values of
the LiH / LiH+ disk fixtures (``tests/test_runners.py:31,69``) and (3) outputs of the
unmodified reference imported in the build container on seeded synthetic inputs
(``tests/golden/make_golden.py`` is the generating script).
"""

from __future__ import annotations

import numpy as np
import scipy.linalg


# --------------------------------------------------------------------------------------
# index maps (utils.py:68-75, utils.py:326-347)
# --------------------------------------------------------------------------------------


def form_vec_energy_differences(moene_occ, moene_virt):
    """ediff[i*nvirt + a] = e_virt[a] - e_occ[i]  (utils.py:326-347)."""
    return (np.asarray(moene_virt)[None, :] - np.asarray(moene_occ)[:, None]).reshape(-1)


def repack_matrix_to_vector(mat):
    """F-order flatten: a (v,o) matrix becomes index i*v + a  (utils.py:68-70)."""
    return np.reshape(mat, -1, order="F")


def form_results(vecs_property, vecs_response):
    """results = P . R^T on (ncomp, dim, 1) stacks  (utils.py:17-27)."""
    assert vecs_property.shape[1:] == vecs_response.shape[1:]
    return np.dot(vecs_property[:, :, 0], vecs_response[:, :, 0].T)


# --------------------------------------------------------------------------------------
# AO -> MO transformation (ao2mo.py:35-70; UHF tuple order interfaces/pyscf/ao2mo.py:49-109)
# --------------------------------------------------------------------------------------


def ao2mo_transform(I, C1, C2, C3, C4):
    """Four quarter transformations in index order 1,2,3,4  (ao2mo.py:44-50)."""
    nao = I.shape[0]
    MO = np.dot(C1.T, I.reshape(nao, -1)).reshape(C1.shape[1], nao, nao, nao)
    MO = np.einsum("qB,Aqrs->ABrs", C2, MO)
    MO = np.einsum("rC,ABrs->ABCs", C3, MO)
    MO = np.einsum("sD,ABCs->ABCD", C4, MO)
    return MO


def ao2mo_rhf_full(I, C):
    """((pq|rs),)  (ao2mo.py:52-56)."""
    C = C[0] if C.ndim == 3 else C
    return (ao2mo_transform(I, C, C, C, C),)


def ao2mo_rhf_partial(I, C, nocc):
    """((ia|jb), (ij|ab))  (ao2mo.py:58-70)."""
    C = C[0] if C.ndim == 3 else C
    Co, Cv = C[:, :nocc], C[:, nocc:]
    return (ao2mo_transform(I, Co, Cv, Co, Cv), ao2mo_transform(I, Co, Co, Cv, Cv))


def ao2mo_uhf_full(I, C):
    """(aaaa, aabb, bbaa, bbbb)  (interfaces/pyscf/ao2mo.py:49-69)."""
    Ca, Cb = C[0], C[1]
    return (
        ao2mo_transform(I, Ca, Ca, Ca, Ca),
        ao2mo_transform(I, Ca, Ca, Cb, Cb),
        ao2mo_transform(I, Cb, Cb, Ca, Ca),
        ao2mo_transform(I, Cb, Cb, Cb, Cb),
    )


def ao2mo_uhf_partial(I, C, occupations):
    """(ovov_aaaa, ovov_aabb, ovov_bbaa, ovov_bbbb, oovv_aaaa, oovv_bbbb)
    (interfaces/pyscf/ao2mo.py:71-109)."""
    na, _, nb, _ = occupations
    Coa, Cva, Cob, Cvb = C[0][:, :na], C[0][:, na:], C[1][:, :nb], C[1][:, nb:]
    return (
        ao2mo_transform(I, Coa, Cva, Coa, Cva),
        ao2mo_transform(I, Coa, Cva, Cob, Cvb),
        ao2mo_transform(I, Cob, Cvb, Coa, Cva),
        ao2mo_transform(I, Cob, Cvb, Cob, Cvb),
        ao2mo_transform(I, Coa, Coa, Cva, Cva),
        ao2mo_transform(I, Cob, Cob, Cvb, Cvb),
    )


# --------------------------------------------------------------------------------------
# Hessian blocks from partially transformed integrals (explicit_equations_partial.py)
# --------------------------------------------------------------------------------------


def _ediff(E_MO, nocc):
    e = np.diag(E_MO)
    return form_vec_energy_differences(e[:nocc], e[nocc:])


def a_singlet_partial(E_MO, ovov, oovv):
    """A = 2(ia|jb) - (ij|ab) + diag  (explicit_equations_partial.py:8-40)."""
    o, v = ovov.shape[0], ovov.shape[1]
    A = 2 * ovov
    A -= oovv.swapaxes(1, 2)
    A = A.reshape(o * v, o * v)
    A += np.diag(_ediff(E_MO, o))
    return A


def a_triplet_partial(E_MO, oovv):
    """A = -(ij|ab) + diag  (explicit_equations_partial.py:43-72)."""
    o, v = oovv.shape[0], oovv.shape[2]
    A = np.zeros((o, v, o, v))
    A -= oovv.swapaxes(1, 2)
    A = A.reshape(o * v, o * v)
    A += np.diag(_ediff(E_MO, o))
    return A


def b_singlet_partial(ovov):
    """B_ref = -(2(ia|jb) - (ib|ja))  (explicit_equations_partial.py:75-96)."""
    o, v = ovov.shape[0], ovov.shape[1]
    B = 2 * ovov
    B -= ovov.swapaxes(1, 3)
    return -B.reshape(o * v, o * v)


def b_triplet_partial(ovov):
    """B_ref = +(ib|ja)  (explicit_equations_partial.py:99-118)."""
    o, v = ovov.shape[0], ovov.shape[1]
    B = np.zeros((o, v, o, v))
    B -= ovov.swapaxes(1, 3)
    return -B.reshape(o * v, o * v)


def a_singlet_ss_partial(E_MO, ovov, oovv):
    """A_ss = (ia|jb) - (ij|ab) + diag  (explicit_equations_partial.py:121-151)."""
    o, v = ovov.shape[0], ovov.shape[1]
    A = ovov.copy()
    A -= oovv.swapaxes(1, 2)
    A = A.reshape(o * v, o * v)
    A += np.diag(_ediff(E_MO, o))
    return A


def a_singlet_os_partial(ovov_xxyy):
    """A_os = (ia|jb)_xxyy  (explicit_equations_partial.py:154-171)."""
    ox, vx, oy, vy = ovov_xxyy.shape
    return ovov_xxyy.copy().reshape(ox * vx, oy * vy)


def b_singlet_ss_partial(ovov):
    """B_ss = -((ia|jb) - (ib|ja))  (explicit_equations_partial.py:174-194)."""
    o, v = ovov.shape[0], ovov.shape[1]
    B = ovov.copy()
    B -= ovov.swapaxes(1, 3)
    return -B.reshape(o * v, o * v)


def b_singlet_os_partial(ovov_xxyy):
    """B_os = -(ia|jb)_xxyy  (explicit_equations_partial.py:197-214)."""
    ox, vx, oy, vy = ovov_xxyy.shape
    return -(ovov_xxyy.copy().reshape(ox * vx, oy * vy))


# --------------------------------------------------------------------------------------
# Hessian blocks from fully transformed integrals (explicit_equations_full.py).
# The reference runs four nested Python loops; the gathers below address the same
# elements (file:line of the element formula is cited per function).
# --------------------------------------------------------------------------------------


def _grids(ox, vx, oy, vy):
    i = np.arange(ox)[:, None, None, None]
    a = np.arange(vx)[None, :, None, None] + ox
    j = np.arange(oy)[None, None, :, None]
    b = np.arange(vy)[None, None, None, :] + oy
    return i, a, j, b


def a_singlet_full(E_MO, TEI, nocc):
    """A[ia,jb] = 2 T[a,i,j,b] - T[a,b,j,i] + diag  (explicit_equations_full.py:28-38)."""
    n = E_MO.shape[0]
    v = n - nocc
    i, a, j, b = _grids(nocc, v, nocc, v)
    A = (2 * TEI[a, i, j, b] - TEI[a, b, j, i]).reshape(nocc * v, nocc * v)
    A = A + np.diag(_ediff(E_MO, nocc))
    return A


def a_triplet_full(E_MO, TEI, nocc):
    """A[ia,jb] = -T[a,b,j,i] + diag  (explicit_equations_full.py:62-70)."""
    n = E_MO.shape[0]
    v = n - nocc
    i, a, j, b = _grids(nocc, v, nocc, v)
    A = (-TEI[a, b, j, i]).reshape(nocc * v, nocc * v)
    A = A + np.diag(_ediff(E_MO, nocc))
    return A


def b_singlet_full(TEI, nocc):
    """B_ref[ia,jb] = -(2 T[a,i,b,j] - T[a,j,b,i])  (explicit_equations_full.py:90-100)."""
    n = TEI.shape[0]
    v = n - nocc
    i, a, j, b = _grids(nocc, v, nocc, v)
    return -((2 * TEI[a, i, b, j] - TEI[a, j, b, i]).reshape(nocc * v, nocc * v))


def b_triplet_full(TEI, nocc):
    """B_ref[ia,jb] = +T[a,j,b,i]  (explicit_equations_full.py:116-124)."""
    n = TEI.shape[0]
    v = n - nocc
    i, a, j, b = _grids(nocc, v, nocc, v)
    return -((-TEI[a, j, b, i]).reshape(nocc * v, nocc * v))


def a_singlet_ss_full(E_MO, TEI, nocc):
    """A_ss[ia,jb] = T[a,i,j,b] - T[a,b,j,i] + diag  (explicit_equations_full.py:145-153)."""
    n = E_MO.shape[0]
    v = n - nocc
    i, a, j, b = _grids(nocc, v, nocc, v)
    A = (TEI[a, i, j, b] - TEI[a, b, j, i]).reshape(nocc * v, nocc * v)
    A = A + np.diag(_ediff(E_MO, nocc))
    return A


def a_singlet_os_full(TEI_xxyy, nocc_x, nocc_y):
    """A_os[ia,jb] = T_xxyy[a,i,j,b]  (explicit_equations_full.py:175-184)."""
    vx = TEI_xxyy.shape[0] - nocc_x
    vy = TEI_xxyy.shape[2] - nocc_y
    i, a, j, b = _grids(nocc_x, vx, nocc_y, vy)
    return np.array(TEI_xxyy[a, i, j, b]).reshape(nocc_x * vx, nocc_y * vy)


def b_singlet_ss_full(TEI, nocc):
    """B_ss[ia,jb] = -(T[a,i,b,j] - T[a,j,b,i])  (explicit_equations_full.py:201-209)."""
    n = TEI.shape[0]
    v = n - nocc
    i, a, j, b = _grids(nocc, v, nocc, v)
    return -((TEI[a, i, b, j] - TEI[a, j, b, i]).reshape(nocc * v, nocc * v))


def b_singlet_os_full(TEI_xxyy, nocc_x, nocc_y):
    """B_os[ia,jb] = -T_xxyy[a,i,b,j]  (explicit_equations_full.py:229-238)."""
    vx = TEI_xxyy.shape[0] - nocc_x
    vy = TEI_xxyy.shape[2] - nocc_y
    i, a, j, b = _grids(nocc_x, vx, nocc_y, vy)
    return -(np.array(TEI_xxyy[a, i, b, j]).reshape(nocc_x * vx, nocc_y * vy))


# --------------------------------------------------------------------------------------
# Operator right-hand sides (operators.py:57-125)
# --------------------------------------------------------------------------------------


def closed_secondary_pairs(occupations):
    """indices.form_indices_from_occupations -> indices_closed_secondary (indices.py:16-37)."""
    na, va, nb, vb = occupations
    norb = na + va
    nact = abs(int(na - nb))
    nclosed = (na + nb - nact) // 2
    nsec = norb - (nclosed + nact)
    return [(i, a) for i in range(nclosed) for a in range(nclosed + nact, nclosed + nact + nsec)]


def form_rhs(ao_integrals, C, occupations, is_imaginary=False, triplet=False, hsofac=None):
    """Property gradient vectors and supervectors  (operators.py:57-125).

    Returns a dict with mo_integrals_ai_alph (ncomp,nov,1), ..._supervector_alph
    (ncomp,2nov,1) and the _beta analogues for UHF.
    """
    if C.ndim == 2:
        C = C[None]
    is_uhf = C.shape[0] == 2
    na, _, nb, _ = occupations
    bpref = +1 if is_imaginary else -1
    out = {}
    for tag, Cs, nocc in (("alph", C[0], na),) + ((("beta", C[1], nb),) if is_uhf else ()):
        vecs, svecs = [], []
        for idx in range(ao_integrals.shape[0]):
            comp = np.dot(Cs[:, nocc:].T, np.dot(ao_integrals[idx], Cs[:, :nocc]))
            if triplet:
                for i, a in closed_secondary_pairs(occupations):
                    comp[a - nocc, i] = 0.0
            vec = repack_matrix_to_vector(comp)[:, None]
            if hsofac is not None:
                vec = vec * hsofac
            vecs.append(vec)
            svecs.append(np.concatenate((vec, vec * bpref), axis=0))
        out[f"mo_integrals_ai_{tag}"] = np.stack(vecs, axis=0)
        out[f"mo_integrals_ai_supervector_{tag}"] = np.stack(svecs, axis=0)
    return out


# --------------------------------------------------------------------------------------
# Explicit super-Hessian (solvers.py:166-385)
# --------------------------------------------------------------------------------------


def _superoverlap(nov, frequency):
    return (
        np.block(
            [[np.eye(nov), np.zeros((nov, nov))], [np.zeros((nov, nov)), -np.eye(nov)]]
        )
        * frequency
    )


def rhf_ab(moenergies, tei_mo, tei_mo_type, nocc, hamiltonian, spin):
    """A and B_ref for an RHF reference  (solvers.py:189-228)."""
    E = moenergies[0]
    nov = nocc * (E.shape[0] - nocc)
    if tei_mo_type == "full":
        (T,) = tei_mo
        A = a_singlet_full(E, T, nocc) if spin == "singlet" else a_triplet_full(E, T, nocc)
        if hamiltonian == "tda":
            B = np.zeros((nov, nov))
        else:
            B = b_singlet_full(T, nocc) if spin == "singlet" else b_triplet_full(T, nocc)
    else:
        ovov, oovv = tei_mo
        A = a_singlet_partial(E, ovov, oovv) if spin == "singlet" else a_triplet_partial(E, oovv)
        if hamiltonian == "tda":
            B = np.zeros((nov, nov))
        else:
            B = b_singlet_partial(ovov) if spin == "singlet" else b_triplet_partial(ovov)
    return A, B


def explicit_hessian_rhf(moenergies, tei_mo, tei_mo_type, nocc, hamiltonian, spin, frequency):
    """G(w) = [[A,B],[B,A]] - w*diag(1,-1)  (solvers.py:179-231)."""
    A, B = rhf_ab(moenergies, tei_mo, tei_mo_type, nocc, hamiltonian, spin)
    nov = A.shape[0]
    return np.block([[A, B], [B, A]]) - _superoverlap(nov, frequency)


def uhf_blocks(moenergies, tei_mo, tei_mo_type, occupations, hamiltonian, spin):
    """The eight ss/os blocks for a UHF reference  (solvers.py:233-370)."""
    na, va, nb, vb = occupations
    nova, novb = na * va, nb * vb
    Ea, Eb = moenergies[0], moenergies[1]
    zab = np.zeros((nova, novb))
    zba = zab.T
    if tei_mo_type == "full":
        aaaa, aabb, bbaa, bbbb = tei_mo
        if spin == "singlet":
            A_ss_a = a_singlet_ss_full(Ea, aaaa, na)
            A_os_a = a_singlet_os_full(aabb, na, nb)
            A_ss_b = a_singlet_ss_full(Eb, bbbb, nb)
            A_os_b = a_singlet_os_full(bbaa, nb, na)
            B_ss_a = b_singlet_ss_full(aaaa, na)
            B_os_a = b_singlet_os_full(aabb, na, nb)
            B_ss_b = b_singlet_ss_full(bbbb, nb)
            B_os_b = b_singlet_os_full(bbaa, nb, na)
        else:
            A_ss_a = a_triplet_full(Ea, aaaa, na)
            A_ss_b = a_triplet_full(Eb, bbbb, nb)
            B_ss_a = b_triplet_full(aaaa, na)
            B_ss_b = b_triplet_full(bbbb, nb)
            A_os_a, B_os_a, A_os_b, B_os_b = zab, zab, zba, zba
    else:
        ovov_aaaa, ovov_aabb, ovov_bbaa, ovov_bbbb, oovv_aaaa, oovv_bbbb = tei_mo
        if spin == "singlet":
            A_ss_a = a_singlet_ss_partial(Ea, ovov_aaaa, oovv_aaaa)
            A_os_a = a_singlet_os_partial(ovov_aabb)
            A_ss_b = a_singlet_ss_partial(Eb, ovov_bbbb, oovv_bbbb)
            A_os_b = a_singlet_os_partial(ovov_bbaa)
            B_ss_a = b_singlet_ss_partial(ovov_aaaa)
            B_os_a = b_singlet_os_partial(ovov_aabb)
            B_ss_b = b_singlet_ss_partial(ovov_bbbb)
            B_os_b = b_singlet_os_partial(ovov_bbaa)
        else:
            A_ss_a = a_triplet_partial(Ea, oovv_aaaa)
            A_ss_b = a_triplet_partial(Eb, oovv_bbbb)
            B_ss_a = b_triplet_partial(ovov_aaaa)
            B_ss_b = b_triplet_partial(ovov_bbbb)
            A_os_a, B_os_a, A_os_b, B_os_b = zab, zab, zba, zba
    if hamiltonian == "tda":
        B_ss_a = np.zeros((nova, nova))
        B_ss_b = np.zeros((novb, novb))
        B_os_a, B_os_b = zab, zba
    return A_ss_a, A_os_a, B_ss_a, B_os_a, A_ss_b, A_os_b, B_ss_b, B_os_b


def explicit_hessian_uhf(moenergies, tei_mo, tei_mo_type, occupations, hamiltonian, spin, frequency):
    """(G_aa - wD, G_ab, G_ba, G_bb - wD)  (solvers.py:372-385)."""
    na, va, nb, vb = occupations
    A_ss_a, A_os_a, B_ss_a, B_os_a, A_ss_b, A_os_b, B_ss_b, B_os_b = uhf_blocks(
        moenergies, tei_mo, tei_mo_type, occupations, hamiltonian, spin
    )
    G_aa = np.block([[A_ss_a, B_ss_a], [B_ss_a, A_ss_a]])
    G_ab = np.block([[A_os_a, B_os_a], [B_os_a, A_os_a]])
    G_ba = np.block([[A_os_b, B_os_b], [B_os_b, A_os_b]])
    G_bb = np.block([[A_ss_b, B_ss_b], [B_ss_b, A_ss_b]])
    return (
        G_aa - _superoverlap(na * va, frequency),
        G_ab,
        G_ba,
        G_bb - _superoverlap(nb * vb, frequency),
    )


# --------------------------------------------------------------------------------------
# Inversion + response vectors (solvers.py:391-466, 499-555)
# --------------------------------------------------------------------------------------


def _inv_plain(G):
    return np.linalg.inv(G)  # solvers.py:494-505


def _inv_cholesky(G):
    fac = scipy.linalg.cholesky(G)  # solvers.py:527-530
    fac_inv = scipy.linalg.inv(fac)
    return np.dot(fac_inv, fac_inv.T)


def response_vectors_rhf(G, supervectors, cholesky=False):
    """rsp_c = G^-1 . rhs_c per component  (solvers.py:401-425)."""
    Ginv = _inv_cholesky(G) if cholesky else _inv_plain(G)
    out = np.empty_like(supervectors)
    for c in range(supervectors.shape[0]):
        out[c] = np.dot(Ginv, supervectors[c])
    return out, Ginv


def response_vectors_uhf(G4, sv_alph, sv_beta, cholesky=False):
    """Block elimination over the spin blocks  (solvers.py:426-466, 506-520, 531-555)."""
    inv = _inv_cholesky if cholesky else _inv_plain
    G_aa, G_ab, G_ba, G_bb = G4
    G_aa_inv, G_bb_inv = inv(G_aa), inv(G_bb)
    left_a = inv(G_aa - np.dot(G_ab, np.dot(G_bb_inv, G_ba)))
    left_b = inv(G_bb - np.dot(G_ba, np.dot(G_aa_inv, G_ab)))
    ra, rb = np.empty_like(sv_alph), np.empty_like(sv_beta)
    for c in range(sv_alph.shape[0]):
        right_a = sv_alph[c] - np.dot(G_ab, np.dot(G_bb_inv, sv_beta[c]))
        right_b = sv_beta[c] - np.dot(G_ba, np.dot(G_aa_inv, sv_alph[c]))
        ra[c] = np.dot(left_a, right_a)
        rb[c] = np.dot(left_b, right_b)
    return ra, rb, (G_aa_inv, G_bb_inv, left_a, left_b)


# --------------------------------------------------------------------------------------
# Result contraction (cphf.py:68-253)
# --------------------------------------------------------------------------------------


def _assemble_blocks(ops_p_alph, ops_r_alph, ops_p_beta=None, ops_r_beta=None):
    dim = sum(p.shape[0] for p in ops_p_alph)
    results = np.zeros((dim, dim))
    rs = 0
    for i1, p1 in enumerate(ops_p_alph):
        cs = 0
        for i2, r2 in enumerate(ops_r_alph):
            blk = form_results(p1, r2)
            if ops_p_beta is not None:
                blk = blk + form_results(ops_p_beta[i1], ops_r_beta[i2])
            results[rs : rs + blk.shape[0], cs : cs + blk.shape[1]] = blk
            cs += r2.shape[0]
        rs += p1.shape[0]
    return results


def coupled_results(sv_alph_list, rsp_alph_list, sv_beta_list=None, rsp_beta_list=None):
    """results = pref * sum_sigma P.R^T, pref = 2 (RHF) / 1 (UHF)  (cphf.py:165-253)."""
    is_uhf = sv_beta_list is not None
    res = _assemble_blocks(sv_alph_list, rsp_alph_list, sv_beta_list, rsp_beta_list)
    return 2 * res / 2 if is_uhf else 4 * res / 2


def uncoupled_results(moenergies, occupations, frequency, sv_alph_list, sv_beta_list=None):
    """results = pref * sum_sigma P.(P/ediff)^T  (cphf.py:68-163)."""
    na, _, nb, _ = occupations
    is_uhf = sv_beta_list is not None

    def _esv(E, nocc):
        e = np.diag(E)
        d = form_vec_energy_differences(e[:nocc], e[nocc:])
        return np.concatenate((d - frequency, d + frequency))[:, None]

    ea = _esv(moenergies[0], na)
    ra = [p / ea for p in sv_alph_list]
    rb = None
    if is_uhf:
        eb = _esv(moenergies[1], nb)
        rb = [p / eb for p in sv_beta_list]
    res = _assemble_blocks(sv_alph_list, ra, sv_beta_list, rb)
    return 2 * res / 2 if is_uhf else 4 * res / 2


# --------------------------------------------------------------------------------------
# Whole-path drivers used by the tests and the CPU baseline
# --------------------------------------------------------------------------------------


def run_cphf(
    C,
    moenergies,
    occupations,
    tei_mo,
    tei_mo_type,
    operators,
    frequencies,
    hamiltonian="rpa",
    spin="singlet",
    cholesky=False,
):
    """CPHF.run with an ExactInv / ExactInvCholesky solver  (cphf.py:56-66, solvers.py:468-480).

    ``operators`` is a list of dicts {ao_integrals, is_imaginary, triplet(optional),
    hsofac(optional)}.  Returns dict(results, uncoupled_results, rspvecs_alph[op][f],
    rspvecs_beta[op][f], rhs[op]).
    """
    if C.ndim == 2:
        C = C[None]
    is_uhf = C.shape[0] == 2
    na, va, nb, vb = occupations
    rhs = [
        form_rhs(
            op["ao_integrals"],
            C,
            occupations,
            is_imaginary=op.get("is_imaginary", False),
            triplet=op.get("triplet", False),
            hsofac=op.get("hsofac"),
        )
        for op in operators
    ]
    rsp_a = [[] for _ in operators]
    rsp_b = [[] for _ in operators]
    for w in frequencies:
        w = float(w)
        if not is_uhf:
            # the reference re-forms A and B at every frequency (solvers.py:477-478)
            G = explicit_hessian_rhf(moenergies, tei_mo, tei_mo_type, na, hamiltonian, spin, w)
            Ginv = _inv_cholesky(G) if cholesky else _inv_plain(G)
            for k, r in enumerate(rhs):
                sv = r["mo_integrals_ai_supervector_alph"]
                out = np.empty_like(sv)
                for c in range(sv.shape[0]):
                    out[c] = np.dot(Ginv, sv[c])
                rsp_a[k].append(out)
        else:
            G4 = explicit_hessian_uhf(
                moenergies, tei_mo, tei_mo_type, occupations, hamiltonian, spin, w
            )
            inv = _inv_cholesky if cholesky else _inv_plain
            G_aa, G_ab, G_ba, G_bb = G4
            G_aa_inv, G_bb_inv = inv(G_aa), inv(G_bb)
            left_a = inv(G_aa - np.dot(G_ab, np.dot(G_bb_inv, G_ba)))
            left_b = inv(G_bb - np.dot(G_ba, np.dot(G_aa_inv, G_ab)))
            for k, r in enumerate(rhs):
                sva = r["mo_integrals_ai_supervector_alph"]
                svb = r["mo_integrals_ai_supervector_beta"]
                oa, ob = np.empty_like(sva), np.empty_like(svb)
                for c in range(sva.shape[0]):
                    right_a = sva[c] - np.dot(G_ab, np.dot(G_bb_inv, svb[c]))
                    right_b = svb[c] - np.dot(G_ba, np.dot(G_aa_inv, sva[c]))
                    oa[c] = np.dot(left_a, right_a)
                    ob[c] = np.dot(left_b, right_b)
                rsp_a[k].append(oa)
                rsp_b[k].append(ob)
    sva_list = [r["mo_integrals_ai_supervector_alph"] for r in rhs]
    svb_list = [r["mo_integrals_ai_supervector_beta"] for r in rhs] if is_uhf else None
    results, unc = [], []
    for f, w in enumerate(frequencies):
        results.append(
            coupled_results(
                sva_list,
                [rsp_a[k][f] for k in range(len(operators))],
                svb_list,
                [rsp_b[k][f] for k in range(len(operators))] if is_uhf else None,
            )
        )
        unc.append(uncoupled_results(moenergies, occupations, float(w), sva_list, svb_list))
    return {
        "results": results,
        "uncoupled_results": unc,
        "rspvecs_alph": rsp_a,
        "rspvecs_beta": rsp_b,
        "rhs": rhs,
    }


# --------------------------------------------------------------------------------------
# Iterative solver with a dense JK (solvers.py:576-660; JK contract integrals.py:41-67,
# interfaces/psi4/integrals.py:143-181).  psi4 is the reference's only working JK and is
# absent here, so the iterative path's parity is anchored on ExactInv property tensors
# (tests/test_solvers.py:108-113 does the same, atol 1e-6).
# --------------------------------------------------------------------------------------


class DenseJK:
    """J_pq = sum (pq|rs) D_rs, K_pq = sum (pr|qs) D_rs with D = C_left . C_right^T."""

    def __init__(self, I):
        self.I = I

    def compute_from_density(self, D):
        J = np.einsum("pqrs,rs->pq", self.I, D)
        K = np.einsum("prqs,rs->pq", self.I, D)
        return J, K

    def compute_from_mocoeffs(self, C_left, C_right=None):
        if C_right is None:
            C_right = C_left
        return self.compute_from_density(np.dot(C_left, C_right.T))


def iterative_rhf(C, moenergies, occupations, ao_integrals, omega, jk, maxiter=40, conv=1e-9):
    """One operator, one frequency of IterativeLinEqSolver._run  (solvers.py:584-660).

    Returns (rspvecs (ncomp,2nov,1) or None if not converged, iterations).
    """
    C = C[0] if C.ndim == 3 else C
    nocc = occupations[0]
    eps = np.diag(moenergies[0])
    ia_denom = eps[nocc:] - eps[:nocc].reshape(-1, 1) - omega
    with np.errstate(divide="ignore", invalid="ignore"):
        all_denom = eps.reshape(-1, 1) - eps - omega
        Co = C[:, :nocc]
        ncomp = ao_integrals.shape[0]
        rhsmats = [C.T @ ao_integrals[c] @ C for c in range(ncomp)]
        x, x_old = [], []
        norb = C.shape[1]
        for rm in rhsmats:
            U = np.zeros((norb, norb))
            U[:nocc, nocc:] = rm[:nocc, nocc:] / ia_denom
            U[nocc:, :nocc] = rm[nocc:, :nocc] / -ia_denom.T
            x.append(U)
            x_old.append(np.zeros_like(U))
        for it in range(1, maxiter + 1):
            for c in range(ncomp):
                J1, K1 = jk.compute_from_mocoeffs(Co, C.dot(x[c].T)[:, :nocc])
                J2, K2 = jk.compute_from_mocoeffs((-C).dot(x[c])[:, :nocc], Co)
                U = x[c].copy()
                upd = (
                    rhsmats[c] - (x[c] * all_denom - (C.T).dot(2 * J1 - K1 + 2 * J2 - K2).dot(C))
                ) / all_denom
                U[:nocc, nocc:] += upd[:nocc, nocc:]
                U[nocc:, :nocc] += upd[nocc:, :nocc]
                x[c] = U.copy()
            rms = []
            for c in range(ncomp):
                rms.append(np.max(x[c] - x_old[c]) ** 2)
                x_old[c] = x[c]
            if max(rms) < conv:
                vec = -np.stack(
                    [
                        np.concatenate(
                            (
                                repack_matrix_to_vector(x[c][:nocc, nocc:].T),
                                repack_matrix_to_vector(x[c][nocc:, :nocc]),
                            )
                        )
                        for c in range(ncomp)
                    ],
                    axis=0,
                )[..., None]
                return vec, it
    return None, maxiter


# --------------------------------------------------------------------------------------
# excitation energies: explicit diagonalisation (solvers.py:663-856) and the transition
# properties built from the eigenvectors (td.py:40-77, :206-241; properties/ecd.py:70-107)
# --------------------------------------------------------------------------------------


def diagonalize_rhf(moenergies, tei_mo, tei_mo_type, nocc, hamiltonian, spin, tda_solver=False):
    """Eigenvalues (ascending, positive half) and unit-norm eigenvectors (columns).

    ``tda_solver=False``: ExactDiagonalizationSolver -- scipy.linalg.eig on [[A, B], [-B, -A]], the
    nov negative roots dropped (solvers.py:758-776).  ``tda_solver=True``:
    ExactDiagonalizationSolverTDA -- scipy.linalg.eig on A alone (solvers.py:841-856)."""
    A, B = rhf_ab(moenergies, tei_mo, tei_mo_type, nocc, hamiltonian, spin)
    nov = A.shape[0]
    if tda_solver:
        assert hamiltonian == "tda"
        eigvals, eigvecs = scipy.linalg.eig(A)
        idx = eigvals.argsort()
        return eigvals[idx], eigvecs[:, idx]
    G = np.block([[A, B], [-B, -A]])
    eigvals, eigvecs = scipy.linalg.eig(G)
    idx = eigvals.argsort()
    return eigvals[idx][nov:], eigvecs[:, idx][:, nov:]


def norm_xy(z, nocc, nvirt):
    """solvers.py:685-690: scale so that 2 (|x|^2 - |y|^2) = 1."""
    x, y = z.reshape(2, nvirt, nocc)
    norm = 1 / np.sqrt(2 * (np.linalg.norm(x) ** 2 - np.linalg.norm(y) ** 2))
    return (x * norm).flatten(), (y * norm).flatten()


def transition_results(eigvals, eigvecs, gradients, nocc, nvirt, tda_driver=False):
    """Per operator: transition moments (nstates, ncomp), oscillator strengths (nstates, ncomp),
    total oscillator strengths (nstates,).

    ``gradients``: list of (ncomp, dim) arrays -- the ai supervector [g; s g] rows for td.TDHF
    (td.py:53-58) or the plain ai rows for td.TDA (td.py:221-226)."""
    nov = nocc * nvirt
    out = [{"transition_moments": [], "oscillator_strengths": [], "total_oscillator_strengths": []}
           for _ in gradients]
    for idx in range(nov):
        eigvec = eigvecs[:, idx].real
        if tda_driver:
            normed = eigvec / np.sqrt(2)
        else:
            x, y = norm_xy(eigvec, nocc, nvirt)
            normed = np.concatenate((x, y))
        e = eigvals[idx].real
        for g, o in zip(gradients, out):
            tm = 2 * np.dot(g, normed)
            o["transition_moments"].append(tm)
            o["oscillator_strengths"].append((2 / 3) * e * tm**2)
            o["total_oscillator_strengths"].append((2 / 3) * e * np.dot(tm, tm))
    return [{k: np.array(v) for k, v in o.items()} for o in out]
