"""Device-resident linear-response engine behind the ExactInv / ExactInvCholesky solvers.

The reference builds the (2 nov)^2 super-Hessian ``G(w) = [[A-w, B],[B, A+w]]`` and calls
``np.linalg.inv`` on it once per frequency (solvers.py:166-231, 468-505).  With
``K = A + B``, ``M = A - B``, ``p = X + Y``, ``m = X - Y`` and a right-hand side ``[g; s g]``
(s = -1 real operator, +1 imaginary; operators.py:75-77) the same equations read

    K p - w m = (1 + s) g            M m - w p = (1 - s) g

so for every frequency   S(w) m = (1 - s) g + w (1 + s) K^-1 g,   S(w) = M - w^2 K^-1
                         K p    = (1 + s) g + w m

with K, M and S(w) symmetric positive definite below the first pole: three Cholesky
factorisations of size nov instead of an LU + explicit inverse of size 2 nov, and ALL
frequencies are factorised in ONE batched call (the GPU analogue of "one solve per frequency",
and what keeps the serial 128-wide panel steps off the critical path).  For UHF the alpha and
beta excitation spaces are stacked; TDA is the special case K = M = A.
When a factorisation reports a non-positive pivot (w above the first pole, unstable reference)
the affected frequencies fall back to LU with partial pivoting on the full G(w).
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from pymolresponse_b200 import _native as nv


def _darr(vals):
    arr = (C.c_double * len(vals))(*[float(v) for v in vals])
    return arr


def potrf(A, batch=1):
    """In-place batched lower Cholesky of (batch, N, N) / (N, N) tensor -> (dinv, info tensor)."""
    L = nv.lib()
    N = int(A.shape[-1])
    lda = int(A.stride(-2))
    stride = int(A.stride(0)) if A.dim() == 3 else 0
    torch = nv.torch_mod()
    dinv = nv.empty(batch, max(1, int(L.pmr_potrf_dinv_doubles(N))))
    info = torch.zeros(batch, dtype=torch.int32, device=nv.device())
    nv.check(L.pmr_potrf(nv.ptr(A), N, lda, batch, stride, nv.ptr(dinv), nv.ptr(info),
                         nv.stream_ptr()))
    return dinv, info


def potrf_rows(A, N, batch=1):
    """Batched lower Cholesky of the leading (N, N) block of (batch, rows, N) matrices whose rows
    N.. hold right-hand sides (one per row): they come out forward-substituted (pmr_potrf_rows)."""
    L = nv.lib()
    rows = int(A.shape[-2])
    assert int(A.shape[-1]) == N and rows >= N
    lda = int(A.stride(-2))
    stride = int(A.stride(0)) if A.dim() == 3 else 0
    torch = nv.torch_mod()
    dinv = nv.empty(batch, max(1, int(L.pmr_potrf_dinv_doubles(N))))
    info = torch.zeros(batch, dtype=torch.int32, device=nv.device())
    nv.check(L.pmr_potrf_rows(nv.ptr(A), N, rows, lda, batch, stride, nv.ptr(dinv), nv.ptr(info),
                              nv.stream_ptr()))
    return dinv, info


def potrs_backward(Lfac, N, dinv, Y, batch=1):
    """L^T X = Y for (batch, N, nrhs) / (N, nrhs) contiguous Y (the forward-substituted right-hand
    sides); returns X in a new tensor of the same shape.  Lfac may carry extra rows (stride only)."""
    L = nv.lib()
    nrhs = int(Y.shape[-1])
    X = nv.empty(*Y.shape)
    if N == 0 or nrhs == 0:
        return X
    lda = int(Lfac.stride(-2))
    sl = int(Lfac.stride(0)) if Lfac.dim() == 3 else 0
    sb = int(X.stride(0)) if X.dim() == 3 else 0
    assert Y.is_contiguous()
    nv.check(L.pmr_potrs_ex(nv.ptr(Lfac), N, lda, batch, sl, nv.ptr(dinv), nv.ptr(X), nrhs,
                            int(X.stride(-2)), sb, nv.ptr(Y), 2, nv.stream_ptr()))
    return X


def symmetrize_lower(A):
    nv.check(nv.lib().pmr_symmetrize_lower(nv.ptr(A), int(A.shape[0]), int(A.stride(0)),
                                           nv.stream_ptr()))
    return A


OUTER_BLOCK = 512  # outer block width of pmr_potrf / pmr_potrf_panel
DISTRIBUTED_POTRF_MIN_N = 4 * OUTER_BLOCK  # below this the redundant single-rank factorisation wins


_comm_streams = {}


def _comm_stream():
    """One side stream per device for the panel broadcasts of :func:`potrf_distributed`."""
    torch = nv.torch_mod()
    dev = torch.cuda.current_device()
    if dev not in _comm_streams:
        _comm_streams[dev] = torch.cuda.Stream(device=dev)
    return _comm_streams[dev]


def _chain_stream():
    """High-priority stream for the panel chain of :func:`potrf_distributed` (the caller's stream
    keeps the follow-up work that may run underneath it)."""
    torch = nv.torch_mod()
    dev = torch.cuda.current_device()
    key = ("chain", dev)
    if key not in _comm_streams:
        _comm_streams[key] = torch.cuda.Stream(device=dev, priority=-1)
    return _comm_streams[key]


def potrf_distributed(A, on_panel=None):
    """Cholesky of ONE replicated (N, N) matrix with the work split over the torch.distributed ranks:
    512-wide block columns are dealt cyclically; the owner factors its block column
    (pmr_potrf_panel) and broadcasts the panel (+ its diagonal-block inverses); every rank stores it
    and applies the rank-512 update to the block columns it owns.  All ranks end up with the complete
    factor, so the triangular solves that follow need no further exchange.  Per-rank flops: u/world
    instead of u.

    Look-ahead: as soon as panel J has arrived, the owner of J+1 updates that one block column,
    factors it and hands it to the broadcast on a side stream; the remaining rank-512 updates of step
    J (the bulk of the flops) run on the main stream underneath the transfer.

    The whole factorisation runs on a high-priority stream; ``on_panel(J, event)`` is called (on the
    caller's stream context) once block column J of the factor and its diagonal-block inverses are
    final on this rank -- work enqueued there on the caller's stream after ``event`` fills the gaps
    of the latency-bound panel chain (the shared K^-1 starts its forward substitutions this way)."""
    from pymolresponse_b200 import distributed as pd

    rank, world = pd.rank_world()
    L = nv.lib()
    torch = nv.torch_mod()
    N = int(A.shape[0])
    lda = int(A.stride(0))
    nbd = int(L.pmr_potrf_dinv_doubles(N))
    dinv = nv.zeros(1, max(1, nbd))
    info = torch.zeros(1, dtype=torch.int32, device=nv.device())
    nblk = (N + OUTER_BLOCK - 1) // OUTER_BLOCK
    caller = torch.cuda.current_stream()
    main = _chain_stream()
    main.wait_stream(caller)
    comm = _comm_stream()
    packs = {}

    def geometry(J):
        j0 = J * OUTER_BLOCK
        jend = min(N, j0 + OUTER_BLOCK)
        kw = jend - j0
        nd = ((kw + 127) // 128) * 128 * 128        # diagonal-block inverses of this panel
        d0 = (j0 // 128) * 128 * 128
        return j0, jend, kw, nd, d0

    def update(J, Jp):
        j0, _, kw, _, _ = geometry(J)
        c0 = Jp * OUTER_BLOCK
        c1 = min(N, c0 + OUTER_BLOCK)
        nv.dgemm(N - c0, c1 - c0, kw, -1.0, A, c0 * lda + j0, lda, 1, A, c0 * lda + j0, 1, lda, 1.0,
                 A, c0 * lda + c0, lda, 1, flags=nv.GEMM_LOWER)

    def post(J):
        """Owner: factor block column J on the main stream; everyone: broadcast it on the side one."""
        j0, jend, kw, nd, d0 = geometry(J)
        owner = J % world
        pack = nv.empty((N - j0) * kw + nd + 1)
        if owner == rank:
            nv.check(L.pmr_potrf_panel(nv.ptr(A), N, lda, j0, nv.ptr(dinv), nv.ptr(info),
                                       nv.stream_ptr()))
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            if owner == rank:
                pack[: (N - j0) * kw].view(N - j0, kw).copy_(A[j0:, j0:jend])
                pack[(N - j0) * kw: (N - j0) * kw + nd].copy_(dinv[0, d0:d0 + nd])
                pack[-1:].copy_(info.to(torch.float64))
            pd.broadcast(pack, owner)
            done = torch.cuda.Event()
            done.record(comm)
        packs[J] = (pack, done)

    with torch.cuda.stream(main):
        post(0)
    for J in range(nblk):
        with torch.cuda.stream(main):
            j0, jend, kw, nd, d0 = geometry(J)
            pack, done = packs.pop(J)
            main.wait_event(done)
            if J % world != rank:
                A[j0:, j0:jend].copy_(pack[: (N - j0) * kw].view(N - j0, kw))
                dinv[0, d0:d0 + nd].copy_(pack[(N - j0) * kw: (N - j0) * kw + nd])
                info.copy_(torch.maximum(info, pack[-1:].to(torch.int32)))
            final = None
            if on_panel is not None:
                final = torch.cuda.Event()
                final.record(main)
            if J + 1 < nblk:
                if (J + 1) % world == rank:
                    update(J, J + 1)
                post(J + 1)
            for Jp in range(J + 2, nblk):
                if Jp % world == rank:
                    update(J, Jp)
            del pack
        if on_panel is not None:
            on_panel(J, final, dinv)
    caller.wait_stream(main)
    return dinv, info


def potrs(Lfac, dinv, B, batch=1, first_row=0, lower_only=False):
    """Solve in place: B is (batch, N, nrhs) or (N, nrhs), contiguous.  ``first_row``: rows of B above
    it are known to be zero (identity columns), the forward substitution starts there; with
    ``lower_only`` the rows of the solution above ``first_row`` are not computed either (left zero)."""
    L = nv.lib()
    N = int(Lfac.shape[-1])
    nrhs = int(B.shape[-1])
    if N == 0 or nrhs == 0:
        return B
    lda = int(Lfac.stride(-2))
    sl = int(Lfac.stride(0)) if Lfac.dim() == 3 else 0
    sb = int(B.stride(0)) if B.dim() == 3 else 0
    work = nv.empty(batch, N, nrhs)
    nv.check(L.pmr_potrs(nv.ptr(Lfac), N, lda, batch, sl, nv.ptr(dinv), nv.ptr(B), nrhs,
                         int(B.stride(-2)), sb, int(first_row), int(bool(lower_only)), nv.ptr(work),
                         nv.stream_ptr()))
    return B


def potri_lower(Lfac, dinv):
    L = nv.lib()
    N = int(Lfac.shape[-1])
    out = nv.empty(N, N)
    work = nv.empty(N, N)
    nv.check(L.pmr_potri_lower(nv.ptr(Lfac), N, int(Lfac.stride(0)), nv.ptr(dinv), nv.ptr(out), N,
                               nv.ptr(work), nv.stream_ptr()))
    return out


def trtri_tree(Lfac, dinv):
    """Z = inv(L) (lower; entries above the 128-wide block diagonal are scratch)."""
    N = int(Lfac.shape[-1])
    Z = nv.empty(N, N)
    nv.check(nv.lib().pmr_trtri_tree(nv.ptr(Lfac), N, int(Lfac.stride(0)), nv.ptr(dinv), nv.ptr(Z), N,
                                     nv.stream_ptr()))
    return Z


def gram_lower_cols(Z, out, c0, ncols):
    """out[c0:, c0:c0+ncols] (lower part) = columns of Z^T Z for lower-triangular Z."""
    N = int(Z.shape[0])
    nv.check(nv.lib().pmr_gram_lower_cols(nv.ptr(Z), N, int(Z.stride(0)), int(c0), int(ncols),
                                          nv.ptr(out), int(out.stride(0)), nv.stream_ptr()))
    return out


#: below this size the shared K^-1 is built from a LOCAL inv(L) (recursive doubling: a handful of
#: launches, N^3/3 flops on every rank) + this rank's columns of Z^T Z; above it the flops of the
#: replicated inv(L) cost more than the latency-bound substitutions they replace
KINV_TREE_MAX_N = 12288


def shifted_combine(M, W, shift_a, shift_b, extra_rows=0):
    """S_f(lower) = M + shift_a[f] I + shift_b[f] W for all f -> (nf, N + extra_rows, N); the extra
    rows (right-hand sides for pmr_potrf_rows) are left for the caller to fill.  The rows are stored
    with an EVEN leading dimension (a view of an (.., N + 1) buffer for odd N, e.g. the UHF joint
    system): the GEMM kernels need 16-byte aligned rows for their TMA / vector paths."""
    L = nv.lib()
    N = int(M.shape[0])
    nf = len(shift_a)
    ld = N + (N & 1)
    S = nv.empty(nf, N + extra_rows, ld)
    nv.check(L.pmr_shifted_combine(nv.ptr(M), nv.ptr(W) if W is not None else C.c_void_p(0), N,
                                   int(M.stride(0)), int(W.stride(0)) if W is not None else N,
                                   _darr(shift_a), _darr(shift_b), nf, nv.ptr(S), ld,
                                   (N + extra_rows) * ld, nv.stream_ptr()))
    return S[:, :, :N] if ld != N else S


def axpby(a, x, b, y):
    """y = a*x + b*y on 2-D tensors whose last dimension is contiguous (row strides free)."""
    rows, cols = int(y.shape[0]), int(y.shape[1])
    assert y.stride(1) == 1 and (x is None or x.stride(1) == 1)
    r0 = 0
    while r0 < rows:  # grid.y limit
        r1 = min(rows, r0 + 65535)
        nv.check(nv.lib().pmr_axpby(r1 - r0, cols, float(a),
                                    nv.ptr(x, r0 * int(x.stride(0))) if x is not None else C.c_void_p(0),
                                    int(x.stride(0)) if x is not None else 0, float(b),
                                    nv.ptr(y, r0 * int(y.stride(0))), int(y.stride(0)),
                                    nv.stream_ptr()))
        r0 = r1
    return y


def getrf(A):
    L = nv.lib()
    torch = nv.torch_mod()
    N = int(A.shape[0])
    ipiv = torch.zeros(max(1, N), dtype=torch.int32, device=nv.device())
    info = torch.zeros(1, dtype=torch.int32, device=nv.device())
    work = nv.empty(int(L.pmr_getrf_work_doubles(N)))
    nv.check(L.pmr_getrf(nv.ptr(A), N, int(A.stride(0)), nv.ptr(ipiv), nv.ptr(info), nv.ptr(work),
                         nv.stream_ptr()))
    return ipiv, info


def getrs(LU, ipiv, B):
    L = nv.lib()
    N, nrhs = int(B.shape[0]), int(B.shape[1])
    work = nv.empty(max(1, N * nrhs + N))   # row buffer + the pivot sequence as one permutation
    nv.check(L.pmr_getrs(nv.ptr(LU), N, int(LU.stride(0)), nv.ptr(ipiv), nv.ptr(B), nrhs,
                         int(B.stride(0)), nv.ptr(work), nv.stream_ptr()))
    return B


def shift_diagonal(A, n, value, offset=0):
    nv.check(nv.lib().pmr_shift_diagonal(nv.ptr(A, offset), int(n), int(A.stride(0)), float(value),
                                         nv.stream_ptr()))


def super_hessian(A, B, frequency):
    """G(w) = [[A, B],[B, A]] - w diag(1, -1) on the device (solvers.py:179-187, 230-231)."""
    N = int(A.shape[0])
    G = nv.empty(2 * N, 2 * N)
    G[:N, :N].copy_(A)
    G[N:, N:].copy_(A)
    if B is None:
        G[:N, N:].zero_()
        G[N:, :N].zero_()
    else:
        G[:N, N:].copy_(B)
        G[N:, :N].copy_(B)
    if frequency != 0.0:
        shift_diagonal(G, N, -frequency, 0)
        shift_diagonal(G, N, +frequency, N * 2 * N + N)
    return G


class LinAlgError(np.linalg.LinAlgError):
    pass


STAGE_TIMES = os.environ.get("PMR_STAGE_TIMES", "0") == "1"


class _StageClock:
    """Optional CUDA-event stopwatch for the solver's stages (PMR_STAGE_TIMES=1 or
    ``_engine.STAGE_TIMES = True``); off by default so that the step has no extra synchronisation."""

    def __init__(self, sink):
        self.on = STAGE_TIMES
        self.sink = sink
        self.marks = []
        self._range = None
        if self.on:
            self.mark("start")

    def stage(self, name):
        """Open the NVTX range ``pmr.<name>`` (closing the previous one): always on, a few hundred
        nanoseconds per call."""
        self._close()
        self._range = nv.nvtx_range(name)
        self._range.__enter__()

    def _close(self):
        if self._range is not None:
            self._range.__exit__(None, None, None)
            self._range = None

    def mark(self, name):
        if self.on:
            e = nv.torch_mod().cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def finish(self):
        self._close()
        if not self.on or len(self.marks) < 2:
            return
        nv.torch_mod().cuda.synchronize()
        ms = self.sink.setdefault("ms", {})
        for (_, e0), (name, e1) in zip(self.marks[:-1], self.marks[1:]):
            ms[name] = ms.get(name, 0.0) + e0.elapsed_time(e1)
        self.marks = []


class HalfSizeSystem:
    """K = A + B_ref and M = A - B_ref on the device; solves all frequencies and columns."""

    def __init__(self, M, K, A=None, B=None, memory_fraction: float = 0.45):
        self.M, self.K = M, K   # lower triangles are what the factorisations read
        self.A, self.B = A, B   # only kept when the LU fallback / explicit G may be needed
        self.N = int(M.shape[0])
        self.memory_fraction = memory_fraction
        self.stats = {"potrf": 0, "potrf_batch": 0, "potri": 0, "lu_fallback": 0}
        # factor of K and K^-1 are kept: later solve() calls and solvers that adopt this system
        # (several properties on one reference wavefunction) reuse them
        self._Kfac = self._dinvK = self._Kinv = None
        self._K_bad = False

    def _chunk(self, nf):
        torch = nv.torch_mod()
        # memory the caching allocator can still hand out, without a driver call
        dev = torch.cuda.current_device()
        total = torch.cuda.get_device_properties(dev).total_memory
        free = total - torch.cuda.memory_allocated(dev) - (2 << 30)
        per = max(1, self.N * self.N * 8)
        return int(max(1, min(nf, (free * self.memory_fraction) // per)))

    def _kinv_blocks(self):
        """Column blocks of the shared K^-1: 2*world blocks, rank r owns blocks r and 2*world-1-r."""
        from pymolresponse_b200 import distributed as pd

        rank, world = pd.rank_world()
        nblocks = 2 * world
        bounds = [pd.shard_range(self.N, b, nblocks) for b in range(nblocks)]
        return rank, world, nblocks, bounds

    def _factor_and_invert_distributed(self, Kf):
        """potrf_distributed(K) with the forward substitutions of this rank's K^-1 column blocks
        following the factorisation panel by panel on the caller's (lower-priority) stream, then the
        backward substitutions and the all-gather.  Returns (dinv, info, Kinv or None)."""
        from pymolresponse_b200 import distributed as pd

        torch = nv.torch_mod()
        L = nv.lib()
        N = self.N
        lda = int(Kf.stride(0))
        rank, world, nblocks, bounds = self._kinv_blocks()
        mine = [bounds[rank], bounds[nblocks - 1 - rank]]
        wmax = max(hi - lo for lo, hi in bounds)
        caller = torch.cuda.current_stream()
        jobs = []
        for slot, (lo, hi) in enumerate(mine):
            w = hi - lo
            if w == 0:
                continue
            rhs = nv.zeros(N, w)
            shift_diagonal(rhs, w, 1.0, offset=lo * w)
            jobs.append({"slot": slot, "w": w, "rhs": rhs, "work": nv.zeros(N, w),
                         "first": (lo // 128) * 128})

        def on_panel(J, final, dinv):
            caller.wait_event(final)
            for job in jobs:
                if (J + 1) * OUTER_BLOCK <= job["first"]:
                    continue                      # this chunk of the right-hand side is still zero
                nv.check(L.pmr_potrs_phase(nv.ptr(Kf), N, lda, nv.ptr(dinv), nv.ptr(job["rhs"]),
                                           job["w"], job["w"], job["first"], 1, nv.ptr(job["work"]),
                                           0, J, J + 1, nv.stream_ptr()))

        dk, info = potrf_distributed(Kf, on_panel=on_panel)
        # no host read of ``info`` here: a failed pivot gives a garbage K^-1 that the caller discards
        # when it looks at all the info words at the end of the solve
        local = nv.zeros(N, 2 * wmax)
        for job in jobs:
            nv.check(L.pmr_potrs_phase(nv.ptr(Kf), N, lda, nv.ptr(dk), nv.ptr(job["rhs"]), job["w"],
                                       job["w"], job["first"], 1, nv.ptr(job["work"]), 1, 0, 1 << 30,
                                       nv.stream_ptr()))
            local[:, job["slot"] * wmax:job["slot"] * wmax + job["w"]].copy_(job["rhs"])
        return dk, info, self._gather_kinv(local, wmax)

    def _gather_kinv(self, local, wmax):
        from pymolresponse_b200 import distributed as pd

        rank, world, nblocks, bounds = self._kinv_blocks()
        N = self.N
        gathered = pd.allgather_columns(local, 2 * wmax * world)   # equal widths per rank
        Kinv = nv.empty(N, N)
        for r in range(world):
            for k, b in enumerate((r, nblocks - 1 - r)):
                lo, hi = bounds[b]
                src0 = r * 2 * wmax + k * wmax
                Kinv[:, lo:hi].copy_(gathered[:, src0:src0 + (hi - lo)])
        return Kinv

    def _kinv_distributed_tree(self, Kfac, dinvK, clock=None):
        """Small-N variant of :meth:`_kinv_distributed`: every rank inverts the factor itself
        (pmr_trtri_tree) and forms only its column blocks of Z^T Z; the blocks -- 2 * world of them,
        cut at multiples of 128, rank r taking blocks r and 2*world-1-r so that the triangular work
        is balanced -- are all-gathered.  Lower triangle only, the rest zero."""
        from pymolresponse_b200 import distributed as pd

        rank, world = pd.rank_world()
        N = self.N
        nblocks = 2 * world
        nb128 = (N + 127) // 128
        cuts = [min(N, 128 * ((nb128 * b) // nblocks)) for b in range(nblocks)] + [N]
        bounds = [(cuts[b], cuts[b + 1]) for b in range(nblocks)]
        wmax = max(hi - lo for lo, hi in bounds)
        mark = clock.mark if clock is not None else (lambda name: None)
        Z = trtri_tree(Kfac, dinvK)
        mark("K_inverse.trtri")
        local = nv.zeros(N, 2 * wmax)
        full = nv.empty(N, N)                      # column blocks land here at their own offsets
        for k, b in enumerate((rank, nblocks - 1 - rank)):
            lo, hi = bounds[b]
            if hi > lo:
                full[:, lo:hi].zero_()
                gram_lower_cols(Z, full, lo, hi - lo)
                local[:, k * wmax:k * wmax + (hi - lo)].copy_(full[:, lo:hi])
        del Z
        mark("K_inverse.gram_columns")
        gathered = pd.allgather_columns(local, 2 * wmax * world)
        mark("K_inverse.allgather")
        for r in range(world):
            for k, b in enumerate((r, nblocks - 1 - r)):
                lo, hi = bounds[b]
                src0 = r * 2 * wmax + k * wmax
                full[:, lo:hi].copy_(gathered[:, src0:src0 + (hi - lo)])
        return full

    def _kinv_distributed(self, Kfac, dinvK, clock=None):
        """K^-1 with the work split over the ranks: every rank solves K X = E_J for its own identity
        columns against the shared factor and the column blocks are all-gathered (N^2 * 8 bytes in
        total over NVLink).  Only the lower triangle of the result is filled (its one consumer, the
        Cholesky of S(w) = M - w^2 K^-1, reads nothing else; the rest is zero): for a column block
        starting at column c both substitutions skip the rows above c, costing (1-c/N)^2 + 1-(c/N)^2
        in units of N^2 w flops, so the columns are cut into 2*world blocks and rank r takes blocks r
        and 2*world-1-r, which sums to the same work on every rank."""
        from pymolresponse_b200 import distributed as pd

        rank, world = pd.rank_world()
        N = self.N
        if N <= KINV_TREE_MAX_N and N >= 128 * 2 * world:
            return self._kinv_distributed_tree(Kfac, dinvK, clock)
        nblocks = 2 * world
        bounds = [pd.shard_range(N, b, nblocks) for b in range(nblocks)]
        mine = [bounds[rank], bounds[nblocks - 1 - rank]]
        wmax = max(hi - lo for lo, hi in bounds)
        local = nv.zeros(N, 2 * wmax)
        for k, (lo, hi) in enumerate(mine):
            w = hi - lo
            if w == 0:
                continue
            rhs = nv.zeros(N, w)
            shift_diagonal(rhs, w, 1.0, offset=lo * w)
            potrs(Kfac, dinvK, rhs, first_row=(lo // 128) * 128, lower_only=True)
            local[:, k * wmax:k * wmax + w].copy_(rhs)
        return self._gather_kinv(local, wmax)

    def solve(self, g_rows, signs, frequencies, distributed=False, global_frequencies=None):
        """g_rows: (ncols, N) CUDA tensor of property gradients g (first half of the supervector);
        signs[c] = -1 (real) / +1 (imaginary).  Returns (xy, bad): xy is the (nf, ncols, 2N) CUDA
        tensor of supervector solutions [X; Y] of G(w) [X;Y] = [g; s g]; ``bad`` is the set of
        frequency indices whose Cholesky factorisation met a non-positive pivot (their rows of xy
        are meaningless: the caller raises or falls back to :meth:`solve_lu`).

        The factorisations run optimistically: the pivot-failure words of K and of every S(w) stay
        on the device and are read ONCE, after everything has been queued."""
        N = self.N
        g_rows = g_rows.contiguous()
        nf, ncols = len(frequencies), int(g_rows.shape[0])
        if ncols > 64:
            # the packing kernels take <= 64 right-hand-side columns a call: solve them in groups
            # (the factor of K and K^-1 are computed once and reused)
            torch = nv.torch_mod()
            parts, bad = [], set()
            for c0 in range(0, ncols, 64):
                xy_g, bad_g = self.solve(g_rows[c0:c0 + 64].contiguous(), signs[c0:c0 + 64], frequencies,
                                         distributed=distributed, global_frequencies=global_frequencies)
                parts.append(xy_g)
                bad |= bad_g
            return torch.cat(parts, dim=1), bad
        xy = nv.zeros(nf, ncols, 2 * N)
        if N == 0 or ncols == 0:
            return xy, set()
        torch = nv.torch_mod()
        signs = [int(s) for s in signs]
        real_cols = [c for c, s in enumerate(signs) if s < 0]
        imag_cols = [c for c, s in enumerate(signs) if s > 0]
        freqs = [float(w) for w in frequencies]
        # with frequencies dealt out to ranks, K^-1 is a collective job: every rank takes part as soon
        # as ANY rank has a dynamic frequency, whatever its own share looks like
        all_freqs = [float(w) for w in (global_frequencies if global_frequencies is not None else freqs)]
        share_kinv = distributed and any(w != 0.0 for w in all_freqs)
        any_dyn = any(w != 0.0 for w in freqs)
        need_K = any_dyn or bool(imag_cols) or share_kinv
        if nf == 0 and not share_kinv:
            return xy, set()
        G = g_rows.t().contiguous()  # (N, ncols)
        clock = _StageClock(self.stats)

        clock.stage("potrf_K")
        infos = []            # (device int32 tensor, what): what = "K" or list of frequency indices
        Kfac, dinvK, Kinv = self._Kfac, self._dinvK, self._Kinv
        new_K = False
        if need_K and not self._K_bad:
            if Kfac is None:
                Kfac = self.K.clone()
                new_K = True
                if share_kinv and self.N >= DISTRIBUTED_POTRF_MIN_N and self.N > KINV_TREE_MAX_N:
                    # large systems -- collective: every rank is here (share_kinv); the factor is
                    # computed once across the ranks and K^-1 comes out of the same pass.  Smaller
                    # ones factor K locally (a distributed panel chain would be latency bound) and
                    # share only the inverse (_kinv_distributed)
                    dinvK, infoK, Kinv = self._factor_and_invert_distributed(Kfac)
                    symmetrize_lower(Kinv)
                    self.stats["potrf_shared"] = self.stats.get("potrf_shared", 0) + 1
                    self.stats["potri_shared"] = self.stats.get("potri_shared", 0) + 1
                else:
                    dinvK, infoK = potrf(Kfac)
                self.stats["potrf"] += 1
                infos.append((infoK, "K"))
            clock.mark("potrf_K")
            clock.stage("K_inverse")
            if (share_kinv or any_dyn) and Kinv is None:
                if share_kinv:
                    Kinv = self._kinv_distributed(Kfac, dinvK, clock)
                    self.stats["potri_shared"] = self.stats.get("potri_shared", 0) + 1
                else:
                    Kinv = potri_lower(Kfac, dinvK)
                    self.stats["potri"] += 1
                # both triangles: K^-1 is then a plain GEMM operand (p = K^-1 r costs one pass over
                # N^2 doubles instead of two latency-bound triangular sweeps over the factor)
                symmetrize_lower(Kinv)
            clock.mark("K_inverse")
        T = None
        if nf > 0 and not self._K_bad and any_dyn and imag_cols:
            T = nv.matmul(Kinv, G[:, imag_cols].contiguous())  # T = K^-1 g for the imaginary columns

        lib = nv.lib()
        null = C.c_void_p(0)
        imag_col_arr = (C.c_int * max(1, ncols))(*[imag_cols.index(c) if c in imag_cols else -1
                                                   for c in range(ncols)])
        if nf > 0 and not self._K_bad:
            step = min(self._chunk(nf), 64)     # the packing kernels take <= 64 frequencies a call
            for f0 in range(0, nf, step):
                fs = list(range(f0, min(nf, f0 + step)))
                ws = [freqs[f] for f in fs]
                nfc = len(fs)
                # all frequencies of the chunk are built and factorised together
                need_m = bool(real_cols) or any(w != 0.0 for w in ws)
                # an even number of columns keeps the rows of the right-hand sides 16-byte aligned
                # (vector loads in the GEMM: 64 x 7168^2 with 3 columns solves in 13.4 ms, with 4 in
                # under 10); the padding column stays zero
                ncp = ncols + (ncols & 1)
                clock.stage("shifted_combine")
                wsub = ws
                if need_m:
                    # S(w) with the right-hand sides appended as extra rows (one kernel packs them):
                    # the factorisation forward-substitutes them on the way (pmr_potrf_rows), only
                    # L^T m = y remains
                    S = shifted_combine(self.M, Kinv if any_dyn else None, [0.0] * nfc,
                                        [-(w * w) for w in ws], extra_rows=ncp)
                    nv.check(lib.pmr_solve_pack_rhs(
                        nv.ptr(g_rows), int(g_rows.stride(0)), nv.ptr(T) if T is not None else null,
                        len(imag_cols), imag_col_arr, ncols, ncp, _darr(wsub), nfc, N,
                        nv.ptr(S, N * int(S.stride(1))), int(S.stride(1)), int(S.stride(0)),
                        nv.stream_ptr()))
                    clock.mark("shifted_combine")
                    clock.stage("potrf_S")
                    dinvS, infoS = potrf_rows(S, N, batch=nfc)
                    clock.mark("potrf_S")
                    clock.stage("solves")
                    self.stats["potrf_batch"] += nfc
                    infos.append((infoS, fs))
                    Y = S[:, N:, :].transpose(1, 2).contiguous()            # (nfc, N, ncp)
                    m_sol = potrs_backward(S, N, dinvS, Y, batch=nfc)       # m_f, (nfc, N, ncp)
                else:
                    S = dinvS = None
                    m_sol = nv.zeros(nfc, N, ncp)
                # p_f = K^-1 ((1+s) g + w m_f), every frequency a block of columns
                need_p = bool(imag_cols) or any(w != 0.0 for w in ws)
                p_all = None
                if need_p:
                    p_rhs = nv.empty(N, nfc * ncols)
                    nv.check(lib.pmr_solve_pack_p(nv.ptr(m_sol), ncp, nv.ptr(g_rows),
                                                  int(g_rows.stride(0)), imag_col_arr, ncols,
                                                  _darr(wsub), nfc, N, nv.ptr(p_rhs), nv.stream_ptr()))
                    if Kinv is not None:
                        p_all = nv.matmul(Kinv, p_rhs)
                    else:
                        p_all = potrs(Kfac, dinvK, p_rhs)
                # xy[f0 : f0 + nfc] is one contiguous (nfc * ncols, 2N) block
                nv.check(lib.pmr_solve_xy(nv.ptr(p_all) if p_all is not None else null, nv.ptr(m_sol),
                                          ncp, ncols, nfc, N, nv.ptr(xy[f0]), nv.stream_ptr()))
                del S, dinvS, m_sol, p_all
                clock.mark("solves")

        # ---- one host read of every pivot-failure word
        clock.stage("info_readback")
        bad = set()
        if self._K_bad:
            bad = set(range(nf))
        elif infos:
            flat = torch.cat([t.reshape(-1) for t, _ in infos]).cpu().numpy()
            pos = 0
            for t, what in infos:
                cnt = int(t.numel())
                vals = flat[pos:pos + cnt]
                pos += cnt
                if what == "K":
                    if int(vals[0]) != 0:
                        self._K_bad = True      # K itself is not SPD: everything goes to LU
                        bad = set(range(nf))
                else:
                    bad.update(f for f, v in zip(what, vals) if int(v) != 0)
        if new_K and not self._K_bad:
            self._Kfac, self._dinvK = Kfac, dinvK
        if not self._K_bad and Kinv is not None:
            self._Kinv = Kinv
        clock.mark("info_readback")
        clock.finish()
        return xy, bad

    def solve_lu(self, g_rows, signs, w):
        """General path: LU with partial pivoting on the full super-Hessian G(w)."""
        assert self.A is not None, "LU fallback needs the A and B blocks"
        N = self.N
        ncols = int(g_rows.shape[0])
        G = super_hessian(self.A, self.B, w)
        ipiv, info = getrf(G)
        if int(info.item()) != 0:
            raise LinAlgError(f"super-Hessian is singular at frequency {w}")
        rhs = nv.zeros(2 * N, ncols)
        Gt = g_rows.t().contiguous()
        axpby(1.0, Gt, 0.0, rhs[:N])
        for c, s in enumerate(signs):
            col = Gt[:, c:c + 1].contiguous()
            tmp = nv.empty(N, 1)
            axpby(float(s), col, 0.0, tmp)
            rhs[N:, c:c + 1] = tmp
        getrs(G, ipiv, rhs)
        return rhs.t().contiguous()
