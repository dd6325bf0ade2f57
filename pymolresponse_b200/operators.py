"""Property operators: AO integrals -> right-hand-side (property gradient) vectors on the GPU.

Drop-in for the reference's ``pymolresponse/operators.py`` ``Operator`` (same keyword-only
constructor, same attributes after ``form_rhs``):

* ``mo_integrals_ai_alph``            (ncomp, nov, 1)   vec(C_v^T O C_o), index ``i*nvirt + a``
* ``mo_integrals_ai_supervector_alph`` (ncomp, 2nov, 1)  ``[g; b*g]``, b = -1 real / +1 imaginary
* the ``_beta`` analogues for unrestricted references
* ``rspvecs_alph`` / ``rspvecs_beta``: lists the solvers append to, one entry per frequency.

The two GEMMs per component (operators.py:85-87) and the packed-index epilogue (F-order repack,
operators.py:94 + utils.py:68-70) are one ``pmr_form_rhs`` call per spin: the second GEMM's
epilogue writes ``i*nvirt + a`` directly through the strided C store.
"""

from __future__ import annotations

import numpy as np

from pymolresponse_b200 import _native as nv

# fine-structure constant (CODATA 2018 / scipy.constants.alpha) for spin-orbit operators
ALPHA_FINE_STRUCTURE = 0.0072973525693


def _alpha():
    try:
        import scipy.constants as spc

        return spc.alpha
    except Exception:  # pragma: no cover - scipy is present in the image
        return ALPHA_FINE_STRUCTURE


class Operator:
    """Property integrals and their representation as CPHF right-hand sides
    (reference operators.py:15-125)."""

    def __init__(self, *, label: str = "", is_imaginary: bool = False,
                 is_spin_dependent: bool = False, triplet: bool = False, slice_idx: int = -1,
                 ao_integrals) -> None:
        self.label = label
        self.is_imaginary = is_imaginary
        self.is_spin_dependent = is_spin_dependent
        self.triplet = triplet
        self.slice_idx = slice_idx
        if "spinorb" in label:
            self.hsofac = (_alpha() ** 2) / 4
        self.frequencies = None
        self.ao_integrals = ao_integrals
        self.rspvecs_alph = []
        self.rspvecs_beta = []
        # device-resident mirrors (torch CUDA tensors) used by the solvers and drivers
        self._dev = {}
        self._dev_src = {}       # host object each device mirror was made from (identity check)
        self._custom = set()     # spins whose supervector the caller replaced after form_rhs
        self._rspvecs_alph_dev = []
        self._rspvecs_beta_dev = []

    def __str__(self) -> str:
        return (
            f'Operator(label="{self.label}", is_imaginary={self.is_imaginary}, '
            f"is_spin_dependent={self.is_spin_dependent}, triplet={self.triplet}, "
            f"slice_idx={self.slice_idx})"
        )

    def calculate_ao_integrals(self) -> None:
        pass

    def form_rhs(self, C, occupations, indices, *, host_copies: bool = True) -> None:
        """Form the right-hand side for CPHF (reference operators.py:57-125)."""
        ao = self.ao_integrals
        assert isinstance(ao, np.ndarray) or nv.is_device_array(ao)
        if len(C.shape) == 2:
            C = C[None]
        assert len(C.shape) == 3
        assert C.shape[0] in (1, 2)
        is_uhf = C.shape[0] == 2
        assert len(occupations) == 4
        nocc_alph, _, nocc_beta, _ = occupations
        sign = +1.0 if self.is_imaginary else -1.0
        scale = float(self.hsofac) if hasattr(self, "hsofac") else 1.0

        O = nv.to_dev(ao)
        ncomp, n = int(O.shape[0]), int(O.shape[1])
        assert O.shape[1] == O.shape[2]
        L = nv.lib()
        for tag, spin, nocc in (("alph", 0, nocc_alph),) + ((("beta", 1, nocc_beta),) if is_uhf else ()):
            Cs = nv.to_dev(C[spin])
            assert int(Cs.shape[0]) == n
            norb = int(Cs.shape[1])
            nvirt = norb - nocc
            nov = nocc * nvirt
            mask = None
            if self.triplet:
                # closed -> secondary excitations are removed (reference operators.py:91-93)
                m = np.zeros(nov, dtype=np.int32)
                for i, a in indices.indices_closed_secondary:
                    if i < nocc and 0 <= a - nocc < nvirt:
                        m[i * nvirt + (a - nocc)] = 1
                mask = nv.to_dev(m, dtype=nv.torch_mod().int32)
            rhs = nv.zeros(ncomp, 2 * nov)
            rhs_ai = nv.zeros(ncomp, nov)
            ws = nv.empty(max(1, ncomp * n * nocc))
            if nov > 0:
                nv.check(L.pmr_form_rhs(nv.ptr(O), ncomp, n, nv.ptr(Cs), int(Cs.stride(0)), nocc,
                                        sign, scale, nv.ptr(mask), nv.ptr(rhs), nv.ptr(rhs_ai),
                                        nv.ptr(ws), nv.stream_ptr()))
            self._dev[f"ai_{tag}"] = rhs_ai
            self._dev[f"super_{tag}"] = rhs
            if host_copies:
                setattr(self, f"mo_integrals_ai_{tag}", nv.to_host(rhs_ai)[:, :, None])
                setattr(self, f"mo_integrals_ai_supervector_{tag}", nv.to_host(rhs)[:, :, None])
            else:
                setattr(self, f"mo_integrals_ai_{tag}", rhs_ai[:, :, None])
                setattr(self, f"mo_integrals_ai_supervector_{tag}", rhs[:, :, None])
            self._dev_src[f"super_{tag}"] = getattr(self, f"mo_integrals_ai_supervector_{tag}")
            self._custom.discard(tag)

    def form_rhs_geometric(self, *args, **kwargs):
        """Nuclear-displacement right-hand sides need psi4 derivative integrals
        (reference operators.py:127-149); out of scope for the GPU hot path."""
        raise NotImplementedError("form_rhs_geometric needs psi4 derivative integrals (out of scope)")

    # ------------------------------------------------------------------ device mirrors
    def supervector_dev(self, tag: str):
        """(ncomp, 2nov) CUDA tensor of the supervectors.  The reference multiplies by whatever
        ``mo_integrals_ai_supervector_*`` holds when the solver runs (solvers.py:407-410), so a
        caller that REASSIGNS the attribute after ``form_rhs`` (as ``form_rhs_geometric`` does,
        operators.py:127-149) gets the new right-hand side: the mirror is re-uploaded when the
        attribute is a different object (in-place edits of the old array are not seen)."""
        key = f"super_{tag}"
        host = getattr(self, f"mo_integrals_ai_supervector_{tag}", None)
        if key not in self._dev or (host is not None and self._dev_src.get(key) is not host):
            t = host if nv.is_device_array(host) else nv.to_dev(host)
            self._dev[key] = t[:, :, 0].contiguous()
            self._dev_src[key] = host
            self._custom.add(tag)
        return self._dev[key]

    def has_standard_supervector(self, tag: str) -> bool:
        """True while the supervector is the ``[g; -+g]`` that ``form_rhs`` built."""
        self.supervector_dev(tag)
        return tag not in self._custom
