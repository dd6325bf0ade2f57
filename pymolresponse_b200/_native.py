"""ctypes binding of libpmr_b200.so (include/pmr_b200.h) plus device-buffer plumbing.

PyTorch is used for device memory, streams and (in ``distributed``) NCCL only; every FLOP of
the hot path runs in the hand-written sm_100a kernels behind this C ABI.  There is NO CPU
fallback: importing the package works anywhere (so host-side logic can be tested on CPU), but
any compute call raises ``PmrUnavailableError`` when the library or a CUDA device is missing.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_LIBPATH = Path(__file__).resolve().parent / "_lib" / "libpmr_b200.so"
_lib = None

c_ll = C.c_longlong
c_dp = C.c_void_p


class PmrUnavailableError(RuntimeError):
    pass


class PmrError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("alpha", C.c_double), ("beta", C.c_double),
        ("A", c_dp), ("a_sm", c_ll), ("a_sk", c_ll),
        ("B", c_dp), ("b_sk", c_ll), ("b_sn", c_ll),
        ("C", c_dp), ("c_sm", c_ll), ("c_sn", c_ll),
        ("batch1", C.c_int), ("batch2", C.c_int),
        ("a_b1", c_ll), ("a_b2", c_ll), ("b_b1", c_ll), ("b_b2", c_ll),
        ("c_b1", c_ll), ("c_b2", c_ll),
        ("flags", C.c_int),
        ("C2", c_dp), ("c2_sm", c_ll), ("c2_sn", c_ll), ("c2_b1", c_ll), ("c2_b2", c_ll),
    ]


class KetDesc(C.Structure):
    _fields_ = [
        ("Co2", c_dp), ("Cv2", c_dp),
        ("ldc2", C.c_int), ("o2", C.c_int), ("v2", C.c_int),
        ("ovov", c_dp),
    ]


class GatherTerm(C.Structure):
    _fields_ = [("ptr", c_dp), ("s_i", c_ll), ("s_a", c_ll), ("s_j", c_ll), ("s_b", c_ll),
                ("coef", C.c_double)]


class AssembleDesc(C.Structure):
    _fields_ = [
        ("no_x", C.c_int), ("nv_x", C.c_int), ("no_y", C.c_int), ("nv_y", C.c_int),
        ("p1", GatherTerm), ("p2", GatherTerm), ("q1", GatherTerm), ("q2", GatherTerm),
        ("eps_occ", c_dp), ("eps_virt", c_dp),
        ("out_mode", C.c_int), ("i_lo", C.c_int), ("i_hi", C.c_int),
        ("out0", c_dp), ("out1", c_dp), ("ld0", c_ll), ("ld1", c_ll),
    ]


GEMM_LOWER = 1
GEMM_KSTART_M = 2
GEMM_KEND_M = 4
GEMM_C_ALIASES_A = 8

# name -> (restype, argtypes); must list every symbol include/pmr_b200.h declares
SYMBOLS = {
    "pmr_version": (C.c_int, []),
    "pmr_last_error": (C.c_char_p, []),
    "pmr_launch_count": (c_ll, []),
    "pmr_reset_launch_count": (None, []),
    "pmr_dgemm": (C.c_int, [C.POINTER(GemmDesc), c_dp]),
    "pmr_profile_enable": (None, [C.c_int]),
    "pmr_profile_fetch": (C.c_int, [C.POINTER(C.c_double), C.c_int]),
    "pmr_ao2mo_transform_workspace": (c_ll, [C.c_int] * 5),
    "pmr_ao2mo_transform": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, C.c_int, c_dp, C.c_int, C.c_int,
                                      c_dp, C.c_int, C.c_int, c_dp, C.c_int, C.c_int, c_dp, c_dp,
                                      c_ll, c_dp]),
    "pmr_ao2mo_partial_workspace": (c_ll, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(KetDesc), C.c_int, C.c_int]),
    "pmr_ao2mo_partial": (C.c_int, [c_dp, C.c_int, C.c_int, C.c_int, c_dp, c_dp, C.c_int, C.c_int,
                                    C.c_int, C.c_int, C.POINTER(KetDesc), c_dp, C.c_int, C.c_int,
                                    c_dp, c_ll, c_dp]),
    "pmr_ao2mo_oovv_tail_workspace": (c_ll, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "pmr_ao2mo_oovv_tail": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, c_dp,
                                      C.c_int, c_dp, c_ll, c_dp]),
    "pmr_assemble": (C.c_int, [C.POINTER(AssembleDesc), c_dp]),
    "pmr_energy_differences": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, c_dp, c_dp]),
    "pmr_potrf_dinv_doubles": (c_ll, [C.c_int]),
    "pmr_set_lookahead": (None, [C.c_int]),
    "pmr_set_lu_cluster_panel": (None, [C.c_int]),
    "pmr_debug_potf2_timing": (C.c_int, [c_dp]),
    "pmr_potrf": (C.c_int, [c_dp, C.c_int, c_ll, C.c_int, c_ll, c_dp, c_dp, c_dp]),
    "pmr_potrf_rows": (C.c_int, [c_dp, C.c_int, C.c_int, c_ll, C.c_int, c_ll, c_dp, c_dp, c_dp]),
    "pmr_potrs_ex": (C.c_int, [c_dp, C.c_int, c_ll, C.c_int, c_ll, c_dp, c_dp, C.c_int, c_ll, c_ll,
                               c_dp, C.c_int, c_dp]),
    "pmr_symmetrize_lower": (C.c_int, [c_dp, C.c_int, c_ll, c_dp]),
    "pmr_potrf_panel": (C.c_int, [c_dp, C.c_int, c_ll, C.c_int, c_dp, c_dp, c_dp]),
    "pmr_potrs": (C.c_int, [c_dp, C.c_int, c_ll, C.c_int, c_ll, c_dp, c_dp, C.c_int, c_ll, c_ll,
                            C.c_int, C.c_int, c_dp, c_dp]),
    "pmr_set_fused_substitution": (None, [C.c_int]),
    "pmr_potrs_phase": (C.c_int, [c_dp, C.c_int, c_ll, c_dp, c_dp, C.c_int, c_ll, C.c_int, C.c_int,
                                  c_dp, C.c_int, C.c_int, C.c_int, c_dp]),
    "pmr_potri_lower": (C.c_int, [c_dp, C.c_int, c_ll, c_dp, c_dp, c_ll, c_dp, c_dp]),
    "pmr_trtri_tree": (C.c_int, [c_dp, C.c_int, c_ll, c_dp, c_dp, c_ll, c_dp]),
    "pmr_gram_lower_cols": (C.c_int, [c_dp, C.c_int, c_ll, C.c_int, C.c_int, c_dp, c_ll, c_dp]),
    "pmr_shifted_combine": (C.c_int, [c_dp, c_dp, C.c_int, c_ll, c_ll, C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.c_int, c_dp, c_ll, c_ll, c_dp]),
    "pmr_shift_diagonal": (C.c_int, [c_dp, C.c_int, c_ll, C.c_double, c_dp]),
    "pmr_cphf_update": (C.c_int, [c_dp, c_dp, c_dp, c_dp, c_dp, C.c_double, C.c_int, C.c_int,
                                  C.c_int, C.c_int, c_dp, c_dp]),
    "pmr_eri_ket_s2_bytes": (c_ll, [c_ll, C.c_int]),
    "pmr_eri_upload_ket_s2": (C.c_int, [c_dp, c_dp, c_ll, C.c_int, c_dp]),
    "pmr_eri_complete_ket_s2": (C.c_int, [c_dp, c_ll, C.c_int, c_dp]),
    "pmr_jacobi_svd": (C.c_int, [c_dp, c_dp, C.c_int, c_ll, C.c_int, C.c_double, c_dp, c_dp, c_dp,
                                 C.POINTER(C.c_int), C.POINTER(C.c_double), c_dp]),
    "pmr_zero_upper": (C.c_int, [c_dp, C.c_int, c_ll, c_dp]),
    "pmr_scale_columns": (C.c_int, [c_dp, C.c_int, C.c_int, c_ll, c_dp, c_dp]),
    "pmr_normalize_columns": (C.c_int, [c_dp, C.c_int, C.c_int, c_ll, c_dp]),
    "pmr_eri_s8_bytes": (c_ll, [C.c_int, C.c_int]),
    "pmr_eri_upload_s8": (C.c_int, [c_dp, c_dp, C.c_int, C.c_int, c_dp]),
    "pmr_eri_complete_s8": (C.c_int, [c_dp, C.c_int, C.c_int, c_dp]),
    "pmr_getrf": (C.c_int, [c_dp, C.c_int, c_ll, c_dp, c_dp, c_dp, c_dp]),
    "pmr_getrf_work_doubles": (c_ll, [C.c_int]),
    "pmr_getrs": (C.c_int, [c_dp, C.c_int, c_ll, c_dp, c_dp, C.c_int, c_ll, c_dp, c_dp]),
    "pmr_form_rhs": (C.c_int, [c_dp, C.c_int, C.c_int, c_dp, C.c_int, C.c_int, C.c_double,
                               C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "pmr_contract_workspace": (c_ll, [C.c_int, C.c_int, c_ll]),
    "pmr_contract_results": (C.c_int, [c_dp, C.c_int, c_ll, c_dp, C.c_int, c_ll, c_ll, C.c_double,
                                       c_dp, c_dp, c_dp]),
    "pmr_pm_to_xy": (C.c_int, [c_dp, c_dp, C.c_int, c_ll, c_ll, c_ll, c_dp, c_ll, c_dp]),
    "pmr_solve_pack_rhs": (C.c_int, [c_dp, c_ll, c_dp, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int,
                                     C.POINTER(C.c_double), C.c_int, c_ll, c_dp, c_ll, c_ll, c_dp]),
    "pmr_solve_pack_p": (C.c_int, [c_dp, C.c_int, c_dp, c_ll, C.POINTER(C.c_int), C.c_int,
                                   C.POINTER(C.c_double), C.c_int, c_ll, c_dp, c_dp]),
    "pmr_solve_xy": (C.c_int, [c_dp, c_dp, C.c_int, C.c_int, C.c_int, c_ll, c_dp, c_dp]),
    "pmr_axpby": (C.c_int, [c_ll, c_ll, C.c_double, c_dp, c_ll, C.c_double, c_dp, c_ll, c_dp]),
    "pmr_uncoupled_divide": (C.c_int, [c_dp, C.c_int, c_ll, c_dp, C.c_double, c_dp, c_dp]),
    "pmr_jk_contract": (C.c_int, [c_dp, C.c_int, c_dp, C.c_int, c_dp, c_dp, c_dp]),
    "pmr_iter_permute_ov_rows": (C.c_int, [c_dp, C.c_int, C.c_int, C.c_int, c_dp, c_dp]),
    "pmr_iter_sigma_combine": (C.c_int, [c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ll, c_ll, C.c_int,
                                         C.c_double, C.c_double, C.c_double, C.c_double, c_dp, c_dp]),
    "pmr_iter_precond": (C.c_int, [c_dp, c_dp, c_dp, c_ll, C.c_int, c_dp, c_dp]),
    "pmr_iter_coldot_workspace": (c_ll, [c_ll, C.c_int]),
    "pmr_iter_coldot": (C.c_int, [c_dp, c_dp, c_ll, C.c_int, c_dp, c_dp, c_dp]),
    "pmr_iter_update_zr": (C.c_int, [c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_ll, C.c_int, c_dp]),
    "pmr_iter_update_d": (C.c_int, [c_dp, c_dp, c_dp, c_dp, c_ll, C.c_int, c_dp]),
}


def library_path() -> Path:
    return Path(os.environ.get("PMR_B200_LIB", str(_LIBPATH)))


def load_library():
    """dlopen libpmr_b200.so and bind every symbol (no CUDA device needed for this)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise PmrUnavailableError(
            f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc); pymolresponse_b200 has no CPU fallback"
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def lib():
    """Library handle for a compute call: requires the .so AND a CUDA device."""
    L = load_library()
    import torch

    if not torch.cuda.is_available():
        raise PmrUnavailableError(
            "no CUDA device: pymolresponse_b200 runs its arithmetic on sm_100a only (no CPU fallback)"
        )
    return L


def check(status: int) -> None:
    if status != 0:
        msg = load_library().pmr_last_error().decode(errors="replace")
        raise PmrError(f"libpmr_b200 error {status}: {msg}")


# ------------------------------------------------------------------------------- tensors
def torch_mod():
    import torch

    return torch


def device():
    torch = torch_mod()
    if not torch.cuda.is_available():
        raise PmrUnavailableError(
            "no CUDA device: pymolresponse_b200 runs its arithmetic on sm_100a only (no CPU fallback)"
        )
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return C.c_void_p(torch_mod().cuda.current_stream().cuda_stream)


def is_device_array(x) -> bool:
    torch = torch_mod()
    return isinstance(x, torch.Tensor) and x.is_cuda


# ---- large pageable host arrays (the AO ERI tensor a drop-in caller hands over as plain NumPy):
# the driver's own staging of a pageable cudaMemcpy runs at ~10 GB/s (one memcpy thread); here the
# array goes through two pinned staging buffers filled by several threads while the previous chunk is
# on the wire, so the copy runs at the PCIe rate of pinned memory
STAGED_UPLOAD_MIN_BYTES = 1 << 30
STAGED_UPLOAD_CHUNK_BYTES = 256 << 20
STAGED_UPLOAD_THREADS = max(1, min(8, (os.cpu_count() or 2) // 2))
_staging = {}


def _upload_staged(arr):
    """Contiguous float64 NumPy array in pageable memory -> CUDA tensor of the same shape."""
    from concurrent.futures import ThreadPoolExecutor

    torch = torch_mod()
    dev = device()
    key = (dev.index, STAGED_UPLOAD_CHUNK_BYTES)
    if key not in _staging:
        bufs = [torch.empty(STAGED_UPLOAD_CHUNK_BYTES // 8, dtype=torch.float64, pin_memory=True)
                for _ in range(2)]
        _staging[key] = (bufs, [b.numpy() for b in bufs], torch.cuda.Stream(device=dev),
                         ThreadPoolExecutor(STAGED_UPLOAD_THREADS))
    bufs, views, side, pool = _staging[key]
    out = torch.empty(arr.shape, dtype=torch.float64, device=dev)
    flat_dst = out.view(-1)
    flat_src = arr.reshape(-1)
    n = flat_src.shape[0]
    chunk = STAGED_UPLOAD_CHUNK_BYTES // 8
    events = [None, None]
    side.wait_stream(torch.cuda.current_stream())
    nthr = STAGED_UPLOAD_THREADS

    def fill(view, off, cnt):
        step = (cnt + nthr - 1) // nthr
        jobs = [(a, min(cnt, a + step)) for a in range(0, cnt, step)]
        list(pool.map(lambda ab: np.copyto(view[ab[0]:ab[1]], flat_src[off + ab[0]:off + ab[1]]), jobs))

    for k, off in enumerate(range(0, n, chunk)):
        cnt = min(chunk, n - off)
        b = k & 1
        if events[b] is not None:
            events[b].synchronize()            # the DMA that last read this staging buffer is done
        fill(views[b], off, cnt)
        with torch.cuda.stream(side):
            flat_dst[off:off + cnt].copy_(bufs[b][:cnt], non_blocking=True)
            events[b] = torch.cuda.Event()
            events[b].record(side)
    torch.cuda.current_stream().wait_stream(side)
    return out


def to_dev(x, dtype=None):
    """NumPy array / torch tensor -> contiguous FP64 (or given dtype) CUDA tensor."""
    torch = torch_mod()
    dtype = dtype or torch.float64
    if isinstance(x, torch.Tensor):
        t = x.to(device=device(), dtype=dtype)
    else:
        arr = np.ascontiguousarray(np.asarray(x))
        if (arr.dtype == np.float64 and dtype == torch.float64 and arr.nbytes >= STAGED_UPLOAD_MIN_BYTES
                and not torch.from_numpy(arr).is_pinned()):
            return _upload_staged(arr)
        t = torch.from_numpy(arr).to(device=device(), dtype=dtype)
    return t.contiguous()


# measured at n = 256 (upload + completion): 0 "row prefix" (a third of the tensor) 208 + 11 ms,
# 1 "box" (a sixth, short runs) 160 + 19 ms, 2 "tail" (the same sixth in long runs) 105 + 18 ms
ERI_S8_LAYOUT = int(os.environ.get("PMR_ERI_S8_LAYOUT", "2"))


def upload_eri_s8(I, layout=None):  # noqa: E741
    """Host (n,n,n,n) AO ERI tensor with 8-fold permutational symmetry -> full CUDA tensor; only a
    unique part (a third or a sixth of the doubles, see include/pmr_b200.h) crosses PCIe, the rest
    is rebuilt in HBM (pmr_eri_upload_s8 / pmr_eri_complete_s8).  The caller vouches for the
    symmetry."""
    torch = torch_mod()
    layout = ERI_S8_LAYOUT if layout is None else int(layout)
    arr = np.ascontiguousarray(np.asarray(I), dtype=np.float64)
    assert arr.ndim == 4 and len(set(arr.shape)) == 1, "the s8 upload needs the whole (n,n,n,n) tensor"
    n = int(arr.shape[0])
    dev = empty(n, n, n, n)
    L = lib()
    check(L.pmr_eri_upload_s8(C.c_void_p(arr.ctypes.data), ptr(dev), n, layout, stream_ptr()))
    check(L.pmr_eri_complete_s8(ptr(dev), n, layout, stream_ptr()))
    torch.cuda.current_stream().synchronize()   # the host buffer may be reused by the caller
    return dev


def upload_eri_slab_s2(I):  # noqa: E741
    """Host slab (..., n, n) of an AO ERI tensor symmetric in its last two indices -> CUDA tensor;
    half of the bytes cross PCIe (pmr_eri_upload_ket_s2 / pmr_eri_complete_ket_s2)."""
    torch = torch_mod()
    arr = np.ascontiguousarray(np.asarray(I), dtype=np.float64)
    assert arr.ndim >= 2 and arr.shape[-1] == arr.shape[-2]
    n = int(arr.shape[-1])
    rows = int(arr.size // (n * n)) if n else 0
    dev = empty(*arr.shape)
    L = lib()
    check(L.pmr_eri_upload_ket_s2(C.c_void_p(arr.ctypes.data), ptr(dev), rows, n, stream_ptr()))
    check(L.pmr_eri_complete_ket_s2(ptr(dev), rows, n, stream_ptr()))
    torch.cuda.current_stream().synchronize()
    return dev


def to_host(t) -> np.ndarray:
    torch = torch_mod()
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def empty(*shape, dtype=None):
    torch = torch_mod()
    return torch.empty(*shape, dtype=dtype or torch.float64, device=device())


def zeros(*shape, dtype=None):
    torch = torch_mod()
    return torch.zeros(*shape, dtype=dtype or torch.float64, device=device())


def ptr(t, offset_elems: int = 0):
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr() + offset_elems * t.element_size())


class nvtx_range:
    """NVTX range around one stage of the hot path (tracing): shows up in
    Nsight Systems / ncu --nvtx captures as ``pmr.<stage>``; a no-op without a CUDA build of torch."""

    def __init__(self, name: str):
        self.name = "pmr." + name
        self.on = False

    def __enter__(self):
        try:
            torch_mod().cuda.nvtx.range_push(self.name)
            self.on = True
        except Exception:
            self.on = False
        return self

    def __exit__(self, *exc):
        if self.on:
            try:
                torch_mod().cuda.nvtx.range_pop()
            except Exception:
                pass
        return False


def launch_count() -> int:
    return int(load_library().pmr_launch_count())


def reset_launch_count() -> None:
    load_library().pmr_reset_launch_count()


# ------------------------------------------------------------------------------- wrappers
def dgemm(M, N, K, alpha, A, a_off, a_sm, a_sk, B, b_off, b_sk, b_sn, beta, Cm, c_off, c_sm, c_sn,
          batch1=1, batch2=1, a_b=(0, 0), b_b=(0, 0), c_b=(0, 0), flags=0):
    g = GemmDesc()
    g.M, g.N, g.K = int(M), int(N), int(K)
    g.alpha, g.beta = float(alpha), float(beta)
    g.A, g.a_sm, g.a_sk = ptr(A, a_off), int(a_sm), int(a_sk)
    g.B, g.b_sk, g.b_sn = ptr(B, b_off), int(b_sk), int(b_sn)
    g.C, g.c_sm, g.c_sn = ptr(Cm, c_off), int(c_sm), int(c_sn)
    g.batch1, g.batch2 = int(batch1), int(batch2)
    g.a_b1, g.a_b2 = int(a_b[0]), int(a_b[1])
    g.b_b1, g.b_b2 = int(b_b[0]), int(b_b[1])
    g.c_b1, g.c_b2 = int(c_b[0]), int(c_b[1])
    g.flags = int(flags)
    check(lib().pmr_dgemm(C.byref(g), stream_ptr()))


def matmul(A, B, transpose_a=False, transpose_b=False, alpha=1.0):
    """Plain 2-D product on the DMMA kernel (A, B contiguous CUDA FP64 tensors)."""
    M, K = (A.shape[1], A.shape[0]) if transpose_a else A.shape
    K2, N = (B.shape[1], B.shape[0]) if transpose_b else B.shape
    assert K == K2
    out = empty(M, N)
    a_sm, a_sk = (1, A.shape[1]) if transpose_a else (A.shape[1], 1)
    b_sk, b_sn = (1, B.shape[1]) if transpose_b else (B.shape[1], 1)
    dgemm(M, N, K, alpha, A, 0, a_sm, a_sk, B, 0, b_sk, b_sn, 0.0, out, 0, N, 1)
    return out


def term(t=None, offset=0, s_i=0, s_a=0, s_j=0, s_b=0, coef=0.0) -> GatherTerm:
    g = GatherTerm()
    g.ptr = ptr(t, offset) if t is not None else C.c_void_p(0)
    g.s_i, g.s_a, g.s_j, g.s_b = int(s_i), int(s_a), int(s_j), int(s_b)
    g.coef = float(coef)
    return g


def assemble(no_x, nv_x, no_y, nv_y, p1, p2, q1, q2, eps_occ, eps_virt, out_mode, out0, out1,
             ld0=None, ld1=None, out0_off=0, out1_off=0, i_lo=0, i_hi=None):
    d = AssembleDesc()
    d.no_x, d.nv_x, d.no_y, d.nv_y = int(no_x), int(nv_x), int(no_y), int(nv_y)
    d.p1, d.p2, d.q1, d.q2 = p1, p2, q1, q2
    d.eps_occ = ptr(eps_occ)
    d.eps_virt = ptr(eps_virt)
    d.out_mode = int(out_mode)
    d.i_lo = int(i_lo)
    d.i_hi = int(no_x if i_hi is None else i_hi)
    d.out0 = ptr(out0, out0_off) if out0 is not None else C.c_void_p(0)
    d.out1 = ptr(out1, out1_off) if out1 is not None else C.c_void_p(0)
    d.ld0 = int(ld0 if ld0 is not None else no_y * nv_y)
    d.ld1 = int(ld1 if ld1 is not None else no_y * nv_y)
    check(lib().pmr_assemble(C.byref(d), stream_ptr()))
