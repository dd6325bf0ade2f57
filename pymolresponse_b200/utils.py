"""Index maps, shape normalisation and fixture readers.

Mirrors the hot-path part of the reference's ``pymolresponse/utils.py`` (same names, same
argument meaning).  The two arithmetic helpers -- ``form_results`` (utils.py:17-27) and
``form_vec_energy_differences`` (utils.py:326-347) -- run on the GPU through libpmr_b200;
the rest is shape/IO plumbing with no arithmetic.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np

from pymolresponse_b200 import _native as nv


# ------------------------------------------------------------------------------ arithmetic
def _as_rows(t):
    """(ncomp, dim, 1) or (ncomp, dim) device tensor -> (ncomp, dim) contiguous view."""
    if t.dim() == 3:
        assert t.shape[2] == 1
        t = t[:, :, 0]
    return t.contiguous()


def contract_rows(P, R, pref: float = 1.0):
    """Device kernel behind form_results: out[a,b] = pref * sum_k P[a,k] R[b,k]."""
    P = _as_rows(P)
    R = _as_rows(R)
    assert P.shape[1] == R.shape[1]
    nP, nR, length = P.shape[0], R.shape[0], P.shape[1]
    L = nv.lib()
    ws = nv.empty(int(L.pmr_contract_workspace(nP, nR, length)))
    out = nv.empty(nP, nR)
    nv.check(L.pmr_contract_results(nv.ptr(P), nP, length, nv.ptr(R), nR, length, length,
                                    float(pref), nv.ptr(out), nv.ptr(ws), nv.stream_ptr()))
    return out


def form_results(vecs_property, vecs_response):
    """All pairwise contractions of property gradients with response vectors
    (reference utils.py:17-27): ``results = P[:, :, 0] . R[:, :, 0]^T``.

    NumPy in -> NumPy out; CUDA tensors in -> CUDA tensor out.
    """
    assert tuple(vecs_property.shape[1:]) == tuple(vecs_response.shape[1:])
    assert len(vecs_property.shape) == 3
    assert vecs_property.shape[2] == 1
    on_dev = nv.is_device_array(vecs_property) and nv.is_device_array(vecs_response)
    out = contract_rows(nv.to_dev(vecs_property), nv.to_dev(vecs_response))
    return out if on_dev else nv.to_host(out)


def form_vec_energy_differences(moene_occ, moene_virt):
    """ediff[i*nvirt + a] = e_virt[a] - e_occ[i]; virtual index fast (reference utils.py:326-347)."""
    on_dev = nv.is_device_array(moene_occ) and nv.is_device_array(moene_virt)
    eo, ev = nv.to_dev(moene_occ), nv.to_dev(moene_virt)
    no, nvirt = int(eo.shape[0]), int(ev.shape[0])
    out = nv.empty(no * nvirt)
    nv.check(nv.lib().pmr_energy_differences(nv.ptr(eo), no, nv.ptr(ev), nvirt, nv.ptr(out),
                                             nv.stream_ptr()))
    return out if on_dev else nv.to_host(out)


# ------------------------------------------------------------------------------ index maps
def repack_matrix_to_vector(mat):
    """(v, o) matrix -> compound index i*v + a (column-major flatten; reference utils.py:68-70)."""
    return np.reshape(mat, -1, order="F")


def repack_vector_to_matrix(vec, shape):
    """Inverse of :func:`repack_matrix_to_vector` (reference utils.py:73-75)."""
    return vec.reshape(shape, order="F")


def form_indices_orbwin(nocc: int, nvirt: int) -> list[tuple[int, int]]:
    """All occupied-virtual index pairs by absolute orbital position (reference utils.py:457-460)."""
    norb = nocc + nvirt
    return [(i, a) for i in range(0, nocc) for a in range(nocc, norb)]


def form_indices_zero(nocc: int, nvirt: int) -> list[tuple[int, int]]:
    """All occupied-virtual index pairs, both counted from zero (reference utils.py:463-465)."""
    return [(i, a) for i in range(nocc) for a in range(nvirt)]


def fix_mocoeffs_shape(mocoeffs):
    """Normalise MO coefficients to (nden, nao, nmo), nden in {1, 2} (reference utils.py:216-237)."""
    if isinstance(mocoeffs, tuple):
        return fix_mocoeffs_shape(np.stack(mocoeffs, axis=0))
    nd = len(mocoeffs.shape)
    assert nd in (2, 3)
    out = mocoeffs[None] if nd == 2 else mocoeffs
    assert len(out.shape) == 3
    return out


def fix_moenergies_shape(moenergies):
    """Normalise MO energies to (nden, nmo, nmo) diagonal matrices (reference utils.py:240-289)."""
    if isinstance(moenergies, tuple):
        return fix_moenergies_shape(np.stack(moenergies, axis=0))
    shape = moenergies.shape
    nd = len(shape)
    assert nd in (1, 2, 3)
    if nd == 1:
        out = np.diag(moenergies)[None]
    elif nd == 2:
        if shape[0] == shape[1]:
            out = moenergies[None]
        else:
            assert shape[0] in (1, 2)
            if shape[0] == 1:
                # kept as in the reference (utils.py:270-273, marked FIXME there)
                out = np.diag(moenergies[:, 0])[None]
            else:
                out = np.stack((np.diag(moenergies[0, :]), np.diag(moenergies[1, :])), axis=0)
    else:
        assert shape[0] in (1, 2)
        assert shape[1] == shape[2]
        out = moenergies
    assert len(out.shape) == 3 and out.shape[0] in (1, 2) and out.shape[1] == out.shape[2]
    return out


# ------------------------------------------------------------------------------ fixture IO
def np_load(filename):
    """np.load that unwraps single-array .npz archives (reference utils.py:30-39)."""
    arr = np.load(filename)
    if isinstance(arr, np.lib.npyio.NpzFile):
        for first in arr.values():
            return first
    return arr


def np_load_2(filename):
    mat = np_load(filename)
    assert len(mat.shape) == 2
    return mat


def parse_int_file_2(filename, dim: int):
    """'mu nu value' triples (1-based, symmetric) -> (dim, dim) matrix (reference utils.py:49-65)."""
    mat = np.zeros((dim, dim))
    with open(filename) as fh:
        for line in fh:
            mu, nu, val = (float(x) for x in line.split())
            mat[int(mu - 1), int(nu - 1)] = mat[int(nu - 1), int(mu - 1)] = val
    return mat


def read_file_occupations(filename):
    """One line of four integers (reference utils.py:114-134)."""
    tokens = Path(filename).read_text().split()
    assert len(tokens) == 4
    return tuple(int(x) for x in tokens)


def _read_dims_then_values(filename, ndim: int):
    with open(filename) as fh:
        dims = tuple(int(x) for x in next(fh).split())
        vals = np.array([float(line) for line in fh], dtype=float)
    assert len(dims) == ndim
    assert vals.size == int(np.prod(dims))
    return vals.reshape(dims)


def read_file_1(filename):
    """libaview-formatted 1-D array: first line = length (reference utils.py:137-143)."""
    return _read_dims_then_values(filename, 1)


def read_file_2(filename):
    return _read_dims_then_values(filename, 2)


def read_file_3(filename):
    return _read_dims_then_values(filename, 3)


def read_file_4(filename):
    return _read_dims_then_values(filename, 4)


# ------------------------------------------------------------------------------ text / small helpers
class Splitter:
    """Cut a fixed-width record into stripped fields (reference utils.py:184-207); trailing empty
    fields are dropped unless ``truncate=False``."""

    def __init__(self, widths) -> None:
        edges = np.concatenate(([0], np.cumsum(list(widths)))).astype(int)
        self.start_indices = [int(x) for x in edges[:-1]]
        self.end_indices = [int(x) for x in edges[1:]]

    def split(self, line: str, truncate: bool = True):
        fields = [line[a:b].strip() for a, b in zip(self.start_indices, self.end_indices)]
        if truncate:
            while len(fields) > 1 and fields[-1] == "":
                fields.pop()
        return fields


def get_reference_value_from_file(filename, hamiltonian: str, spin: str, frequency: str,
                                  label_1: str, label_2: str) -> float:
    """Look up '<hamiltonian> <spin> <frequency> <label_1> <label_2> <value>' in a reference-value
    file; the frequency is matched as text, the last match wins (reference utils.py:78-111)."""
    wanted = (hamiltonian, spin, frequency, label_1, label_2)
    value = None
    with open(filename) as fh:
        for line in fh:
            tokens = line.split()
            assert len(tokens) == 6
            if tuple(tokens[:5]) == wanted:
                value = float(tokens[5])
    if value is None:
        raise ValueError("Could not find reference value")
    return value


def read_dalton_propfile(tmpdir):
    """Rows of a DALTON.PROP file as lists of fields (reference utils.py:292-302)."""
    splitter = Splitter([5, 3, 4, 11, 23, 9, 9, 9, 9, 23, 23, 23, 4, 4, 4])
    with open(Path(tmpdir) / "DALTON.PROP") as fh:
        return [splitter.split(line) for line in fh.readlines()]


def tensor_printer(tensor):
    """Print a 2-D tensor, its eigenvalues and their mean; returns (eigvals, iso, 0.0) as the
    reference does (utils.py:305-323; the anisotropy is not implemented there either)."""
    print(tensor)
    eigvals = np.linalg.eigvals(tensor)
    iso = np.average(eigvals)
    print(eigvals)
    print(iso)
    return eigvals, iso, 0.0


def screen(mat, thresh: float = 1.0e-16):
    """Copy of ``mat`` with every |element| <= thresh set to zero (reference utils.py:350-370)."""
    out = mat.copy()
    out[np.abs(mat) <= thresh] = 0.0
    return out


def matsym(amat, thrzer: float = 1.0e-14) -> int:
    """1 symmetric, 2 antisymmetric, 3 all (pair sums and differences) below ``thrzer``, 0 neither
    (DALTON's MATSYM convention; reference utils.py:373-411)."""
    assert amat.shape[0] == amat.shape[1]
    symmetric = bool((np.abs(amat - amat.T) <= thrzer).all())
    antisymmetric = bool((np.abs(amat + amat.T) <= thrzer).all())
    return (1 if symmetric else 0) + (2 if antisymmetric else 0)


def flip_triangle_sign(A, triangle: str = "lower"):
    """Copy of a square matrix with the sign of one triangle (diagonal included) flipped
    (reference utils.py:414-442)."""
    assert len(A.shape) == 2 and A.shape[0] == A.shape[1]
    if triangle == "lower":
        mask = np.tril(np.ones(A.shape, dtype=bool))
    elif triangle == "upper":
        mask = np.triu(np.ones(A.shape, dtype=bool))
    else:
        raise ValueError("argument to triangle must be 'upper' or 'lower'")
    return np.where(mask, -A, A)
