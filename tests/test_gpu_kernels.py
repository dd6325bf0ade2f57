"""GPU: each CUDA kernel family of libpmr_b200 against NumPy / the oracle, through the C ABI."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _nv():
    from pymolresponse_b200 import _native as nv

    return nv


def _rng(seed=0):
    return np.random.default_rng(seed)


# ----------------------------------------------------------------------------------- GEMM
def _gemm_case(M, N, K, a_kc, b_kc, pad_a=0, pad_b=0, pad_c=0, alpha=1.0, beta=0.0, seed=0,
               flags=0, offset=0):
    """Build operands in the requested memory layout, run pmr_dgemm, return (got, want)."""
    nv = _nv()
    r = _rng(seed)
    A = r.standard_normal((M, K))
    B = r.standard_normal((K, N))
    C0 = r.standard_normal((M, N))
    if a_kc:   # A stored [m][k]
        Ast = np.zeros((M, K + pad_a)); Ast[:, :K] = A
        a_sm, a_sk = K + pad_a, 1
    else:      # A stored [k][m]
        Ast = np.zeros((K, M + pad_a)); Ast[:, :M] = A.T
        a_sm, a_sk = 1, M + pad_a
    if b_kc:   # B stored [n][k]
        Bst = np.zeros((N, K + pad_b)); Bst[:, :K] = B.T
        b_sk, b_sn = 1, K + pad_b
    else:      # B stored [k][n]
        Bst = np.zeros((K, N + pad_b)); Bst[:, :N] = B
        b_sk, b_sn = N + pad_b, 1
    Cst = np.zeros((M, N + pad_c)); Cst[:, :N] = C0
    # `offset` shifts the base pointers by an odd number of doubles => 8-byte-aligned code path
    Ad = nv.to_dev(np.concatenate((np.zeros(offset), Ast.ravel())))
    Bd = nv.to_dev(np.concatenate((np.zeros(offset), Bst.ravel())))
    Cd = nv.to_dev(np.concatenate((np.zeros(offset), Cst.ravel())))
    nv.dgemm(M, N, K, alpha, Ad, offset, a_sm, a_sk, Bd, offset, b_sk, b_sn, beta, Cd, offset,
             N + pad_c, 1, flags=flags)
    got = nv.to_host(Cd)[offset:].reshape(M, N + pad_c)[:, :N]
    want = alpha * (A @ B) + beta * C0
    return got, want, C0


@pytest.mark.parametrize("a_kc", [True, False])
@pytest.mark.parametrize("b_kc", [True, False])
@pytest.mark.parametrize("shape", [(128, 128, 64), (131, 77, 53), (5, 3, 1), (300, 257, 129),
                                   (64, 1000, 40), (257, 33, 16), (1, 1, 7), (200, 40, 0)])
def test_dgemm_layouts_and_ragged_edges(a_kc, b_kc, shape):
    M, N, K = shape
    got, want, _ = _gemm_case(M, N, K, a_kc, b_kc, alpha=0.75, beta=-0.5)
    scale = max(1.0, np.abs(want).max())
    assert np.abs(got - want).max() <= 1e-13 * scale * max(1, K) ** 0.5


@pytest.mark.parametrize("pads", [(1, 1, 1), (2, 0, 3), (0, 3, 0)])
@pytest.mark.parametrize("offset", [0, 1])
def test_dgemm_unaligned_strides(pads, offset):
    """Odd leading dimensions / odd base offsets force the 8-byte cp.async path."""
    for a_kc in (True, False):
        for b_kc in (True, False):
            got, want, _ = _gemm_case(150, 90, 70, a_kc, b_kc, *pads, alpha=1.0, beta=1.0,
                                      offset=offset, seed=3)
            assert np.abs(got - want).max() <= 1e-12


@pytest.mark.parametrize("a_kc", [True, False])
@pytest.mark.parametrize("b_kc", [True, False])
@pytest.mark.parametrize("shape", [(1000, 96, 256), (130, 34, 64), (4096, 32, 512), (258, 200, 16),
                                   (131, 77, 48)])
def test_dgemm_bulk_copy_path(a_kc, b_kc, shape, monkeypatch):
    """Aligned rows and K % 16 == 0 take the cp.async.bulk + mbarrier (TMA bulk) producer/consumer
    kernel; ragged M/N tiles included (odd extents only for k-contiguous operands)."""
    M, N, K = shape
    if (not a_kc and M % 2) or (not b_kc and N % 2):
        pytest.skip("odd extent of an mn-contiguous operand uses the cp.async kernel (covered elsewhere)")
    got, want, _ = _gemm_case(M, N, K, a_kc, b_kc, alpha=-1.0, beta=1.0, seed=31)
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
    got, want, _ = _gemm_case(M, N, K, a_kc, b_kc, alpha=2.0, beta=0.0, seed=32)
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("shape", [(1000, 96, 256), (130, 34, 70), (4096, 32, 512), (258, 200, 18),
                                   (131, 77, 48), (640, 640, 512), (129, 129, 2)])
def test_dgemm_tma_tensor_map_path(shape):
    """Both operands k-contiguous with even leading dimensions: the cp.async.bulk.tensor (tensor-map
    TMA, 128-byte swizzle) producer/consumer kernel; ragged M / N / K edges are zero-filled by the
    TMA unit.  Compared with the cp.async ring (PMR_GEMM_NO_TMA is read once per process, so the
    cross-check is against NumPy)."""
    M, N, K = shape
    for pad in (0, 2, 6):
        got, want, _ = _gemm_case(M, N, K, True, True, pad_a=pad, pad_b=pad, alpha=-1.0, beta=1.0, seed=41)
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
    got, want, _ = _gemm_case(M, N, K, True, True, alpha=2.0, beta=0.0, seed=42)
    assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())


def test_dgemm_tma_batched_shared_operand_and_lower():
    """Tensor-map path with a two-level batch, one operand shared by the batch (stride 0 -> extent-1
    dimension of the map) and the lower-triangular store of the Cholesky trailing update."""
    nv = _nv()
    r = _rng(17)
    b1, b2, M, K = 3, 2, 300, 64
    P = r.standard_normal((b1, b2, M, K))
    Q = r.standard_normal((b2, M, K))            # shared across b1
    C0 = r.standard_normal((b1, b2, M, M))
    Pd, Qd, Cd = nv.to_dev(P), nv.to_dev(Q), nv.to_dev(C0).clone()
    nv.dgemm(M, M, K, -1.0, Pd, 0, K, 1, Qd, 0, 1, K, 1.0, Cd, 0, M, 1, batch1=b1, batch2=b2,
             a_b=(b2 * M * K, M * K), b_b=(0, M * K), c_b=(b2 * M * M, M * M), flags=nv.GEMM_LOWER)
    want = C0 - np.einsum("xymk,ynk->xymn", P, Q)
    got = nv.to_host(Cd)
    low = np.tril(np.ones((M, M), dtype=bool))
    assert np.abs(got[..., low] - want[..., low]).max() <= 1e-12 * np.abs(want).max()
    assert np.array_equal(got[..., ~low], C0[..., ~low])
    # sub-matrix views of one big matrix (offsets, leading dimension of the parent), as in potrf
    N = 700
    S = r.standard_normal((N, N))
    Sd = nv.to_dev(S).clone()
    r0, j0, kw = 256, 128, 128
    rows = N - r0
    nv.dgemm(rows, rows, kw, -1.0, Sd, r0 * N + j0, N, 1, Sd, r0 * N + j0, 1, N, 1.0,
             Sd, r0 * N + r0, N, 1, flags=nv.GEMM_LOWER)
    want = S.copy()
    upd = S[r0:, r0:] - S[r0:, j0:j0 + kw] @ S[r0:, j0:j0 + kw].T
    lowr = np.tril(np.ones((rows, rows), dtype=bool))
    want[r0:, r0:][lowr] = upd[lowr]
    assert np.abs(nv.to_host(Sd) - want).max() <= 1e-12 * np.abs(want).max()


def test_dgemm_lower_flag_leaves_upper_untouched():
    nv = _nv()
    got, want, C0 = _gemm_case(260, 260, 48, True, True, alpha=-1.0, beta=1.0,
                               flags=nv.GEMM_LOWER, seed=5)
    low = np.tril(np.ones((260, 260), dtype=bool))
    assert np.abs(got[low] - want[low]).max() <= 1e-12
    assert np.array_equal(got[~low], C0[~low])


def test_dgemm_batched_two_levels():
    nv = _nv()
    r = _rng(7)
    b1, b2, M, N, K = 3, 4, 70, 45, 33
    A = r.standard_normal((b1, b2, M, K))
    B = r.standard_normal((b2, K, N))          # shared across b1 (stride 0)
    Ad, Bd = nv.to_dev(A), nv.to_dev(B)
    Cd = nv.zeros(b1, b2, M, N)
    nv.dgemm(M, N, K, 1.0, Ad, 0, K, 1, Bd, 0, N, 1, 0.0, Cd, 0, N, 1, batch1=b1, batch2=b2,
             a_b=(b2 * M * K, M * K), b_b=(0, K * N), c_b=(b2 * M * N, M * N))
    want = np.einsum("xymk,ykn->xymn", A, B)
    assert np.abs(nv.to_host(Cd) - want).max() <= 1e-12


def test_dgemm_strided_c_store_transposes():
    """C written with (row, col) strides (1, M): the packed-index epilogue used by form_rhs."""
    nv = _nv()
    r = _rng(9)
    M, N, K = 37, 19, 23
    A, B = r.standard_normal((M, K)), r.standard_normal((K, N))
    Cd = nv.zeros(N, M)
    nv.dgemm(M, N, K, 1.0, nv.to_dev(A), 0, K, 1, nv.to_dev(B), 0, N, 1, 0.0, Cd, 0, 1, M)
    assert np.abs(nv.to_host(Cd) - (A @ B).T).max() <= 1e-12


def test_dgemm_rejects_bad_strides():
    nv = _nv()
    x = nv.zeros(16)
    with pytest.raises(nv.PmrError):
        nv.dgemm(2, 2, 2, 1.0, x, 0, 3, 2, x, 0, 2, 1, 0.0, x, 0, 2, 1)


# ----------------------------------------------------------------------------------- Cholesky
def _spd(n, seed, batch=None):
    r = _rng(seed)
    shape = (n, n) if batch is None else (batch, n, n)
    X = r.standard_normal(shape)
    return X @ np.swapaxes(X, -1, -2) / max(n, 1) + np.eye(n) * 1.5


@pytest.mark.parametrize("n", [1, 5, 128, 129, 300, 515])
def test_potrf_potrs_batched(n):
    from pymolresponse_b200 import _engine as eng

    nv = _nv()
    batch = 3
    A = _spd(n, 11, batch)
    Ad = nv.to_dev(A).clone()
    dinv, info = eng.potrf(Ad, batch=batch)
    assert (nv.to_host(info) == 0).all()
    L = np.tril(nv.to_host(Ad))
    want = np.linalg.cholesky(A)
    assert np.abs(L - want).max() <= 1e-12 * max(1, n)
    # the strictly upper triangle is neither read nor written
    iu = np.triu_indices(n, 1)
    assert np.array_equal(nv.to_host(Ad)[:, iu[0], iu[1]], A[:, iu[0], iu[1]])
    rhs = _rng(12).standard_normal((batch, n, 4))
    X = nv.to_host(eng.potrs(Ad, dinv, nv.to_dev(rhs).clone(), batch=batch))
    assert np.abs(X - np.linalg.solve(A, rhs)).max() <= 1e-11 * max(1, n)


@pytest.mark.parametrize("n,batch", [(1100, 2), (2177, 1), (2560, 3)])
def test_potrf_lookahead_path(n, batch):
    """N > 1024 takes the two-stream look-ahead path (outer blocks of 512, panels of 128,
    in-place panel solves); residual check, since NumPy's factor is not needed at this size."""
    from pymolresponse_b200 import _engine as eng

    nv = _nv()
    A = _spd(n, 21, batch)
    Ad = nv.to_dev(A).clone()
    dinv, info = eng.potrf(Ad, batch=batch)
    assert (nv.to_host(info) == 0).all()
    L = np.tril(nv.to_host(Ad))
    resid = np.abs(L @ np.swapaxes(L, -1, -2) - A).max()
    assert resid <= 1e-12 * n
    rhs = _rng(22).standard_normal((batch, n, 2))
    X = nv.to_host(eng.potrs(Ad, dinv, nv.to_dev(rhs).clone(), batch=batch))
    assert np.abs(A @ X - rhs).max() <= 1e-10 * n


@pytest.mark.parametrize("n,lo,w", [(700, 0, 90), (700, 300, 77), (700, 640, 60), (515, 512, 3)])
def test_potrs_identity_columns_skip_zero_rows(n, lo, w):
    """potrs(first_row=...) on identity columns lo:lo+w skips the zero block rows of the forward
    substitution; the result equals the general solve bit for bit (the skipped work adds zeros)."""
    from pymolresponse_b200 import _engine as eng

    nv = _nv()
    A = _spd(n, 31)
    Ad = nv.to_dev(A).clone()
    dinv, info = eng.potrf(Ad)
    assert (nv.to_host(info) == 0).all()
    E = np.zeros((n, w))
    E[lo:lo + w] = np.eye(w)
    full = nv.to_host(eng.potrs(Ad, dinv, nv.to_dev(E).clone()))
    fast = nv.to_host(eng.potrs(Ad, dinv, nv.to_dev(E).clone(), first_row=(lo // 128) * 128))
    assert np.array_equal(full, fast)
    assert np.abs(fast - np.linalg.inv(A)[:, lo:lo + w]).max() <= 1e-11 * n
    # lower_only: the rows above the block are left zero, the rest is unchanged
    r0 = (lo // 128) * 128
    low = nv.to_host(eng.potrs(Ad, dinv, nv.to_dev(E).clone(), first_row=r0, lower_only=True))
    assert np.array_equal(low[r0:], full[r0:])
    assert not low[:r0].any()


def test_dgemm_in_place_panel_update():
    """A <- A * B with the aliasing flag (used by the panel solve of the Cholesky)."""
    nv = _nv()
    r = _rng(23)
    M, K = 1000, 128
    A, B = r.standard_normal((M, K)), r.standard_normal((K, K))
    Ad = nv.to_dev(A).clone()
    Bd = nv.to_dev(B.T.copy())          # B stored [n][k]
    nv.dgemm(M, K, K, 1.0, Ad, 0, K, 1, Bd, 0, 1, K, 0.0, Ad, 0, K, 1, flags=nv.GEMM_C_ALIASES_A)
    assert np.abs(nv.to_host(Ad) - A @ B).max() <= 1e-11


def test_potrf_reports_first_bad_pivot():
    from pymolresponse_b200 import _engine as eng

    nv = _nv()
    A = _spd(200, 13)
    A[150, 150] = -5.0
    _, info = eng.potrf(nv.to_dev(A).clone())
    assert int(nv.to_host(info)[0]) == 151
    B = np.stack((_spd(140, 14), -np.eye(140)))
    _, info = eng.potrf(nv.to_dev(B).clone(), batch=2)
    assert list(nv.to_host(info)) == [0, 1]


@pytest.mark.parametrize("n", [3, 128, 200, 385, 640, 1537, 2560, 3000])
def test_potri_lower(n):
    from pymolresponse_b200 import _engine as eng

    nv = _nv()
    A = _spd(n, 15)
    Ad = nv.to_dev(A).clone()
    dinv, info = eng.potrf(Ad)
    W = nv.to_host(eng.potri_lower(Ad, dinv))
    want = np.linalg.inv(A)
    il = np.tril_indices(n)
    assert np.abs(W[il] - want[il]).max() <= 1e-11 * max(1, n)
    # the recursive-doubling triangular inverse on its own (ragged pair sizes at several levels)
    Z = nv.to_host(eng.trtri_tree(Ad, dinv))
    Lref = np.linalg.cholesky(A)
    Zref = np.linalg.inv(Lref)
    assert np.abs(np.tril(Z) - Zref).max() <= 1e-11 * max(1, n)
    for k0 in range(0, n, 128):          # the block diagonal is exact, zeros above the diagonal included
        blk = Z[k0:k0 + 128, k0:k0 + 128]
        assert np.array_equal(np.triu(blk, 1), np.zeros_like(blk))
    # column blocks of Z^T Z (what the ranks share out)
    if n >= 385:
        c0 = 128 * ((n // 128) // 2)
        out = nv.zeros(n, n)
        eng.gram_lower_cols(nv.to_dev(Z), out, c0, min(200, n - c0))
        got = nv.to_host(out)
        w = min(200, n - c0)
        ref = np.tril(want)[:, c0:c0 + w]
        assert np.abs(got[c0:, c0:c0 + w] - ref[c0:]).max() <= 1e-11 * n
        assert np.abs(got[:c0]).max() == 0.0


def test_shifted_combine_lower_triangle():
    from pymolresponse_b200 import _engine as eng

    nv = _nv()
    r = _rng(16)
    n = 70
    M, W = r.standard_normal((n, n)), r.standard_normal((n, n))
    S = nv.to_host(eng.shifted_combine(nv.to_dev(M), nv.to_dev(W), [0.0, 0.25], [-0.5, 2.0]))
    il = np.tril_indices(n)
    for f, (a, b) in enumerate(((0.0, -0.5), (0.25, 2.0))):
        want = M + a * np.eye(n) + b * W
        assert np.abs(S[f][il] - want[il]).max() <= 1e-14


# ----------------------------------------------------------------------------------- LU
@pytest.fixture(params=[1, 0], ids=["cluster_panel", "column_launches"])
def lu_panel_mode(request):
    nv = _nv()
    nv.lib().pmr_set_lu_cluster_panel(request.param)
    yield request.param
    nv.lib().pmr_set_lu_cluster_panel(1)


@pytest.mark.parametrize("n", [1, 7, 64, 65, 200, 256, 257, 333, 700, 1500])
def test_getrf_getrs_with_pivoting(n, lu_panel_mode):
    from pymolresponse_b200 import _engine as eng

    nv = _nv()
    r = _rng(17)
    A = r.standard_normal((n, n))
    if n > 2:
        A[0, 0] = 0.0          # forces a row interchange in the first column
        A[2, 2] = 1e-14
    rhs = r.standard_normal((n, 3 if n != 333 else 200))
    Ad = nv.to_dev(A).clone()
    ipiv, info = eng.getrf(Ad)
    assert int(nv.to_host(info)[0]) == 0
    X = nv.to_host(eng.getrs(Ad, ipiv, nv.to_dev(rhs).clone()))
    want = np.linalg.solve(A, rhs)
    assert np.abs(X - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
    # P A = L U
    LU = nv.to_host(Ad)
    Lm, Um = np.tril(LU, -1) + np.eye(n), np.triu(LU)
    PA = A.copy()
    for j, p in enumerate(nv.to_host(ipiv)[:n]):
        PA[[j, p]] = PA[[p, j]]
    assert np.abs(Lm @ Um - PA).max() <= 1e-12 * max(1, n)
    # partial pivoting: every multiplier is bounded by one, and the pivots are LAPACK's
    assert np.abs(np.tril(LU, -1)).max() <= 1.0
    import scipy.linalg

    _, piv = scipy.linalg.lu_factor(A)
    assert np.array_equal(nv.to_host(ipiv)[:n], piv)


def test_getrf_panel_modes_agree_bitwise():
    from pymolresponse_b200 import _engine as eng

    nv = _nv()
    A = _rng(171).standard_normal((777, 777))
    out = []
    for mode in (1, 0):
        nv.lib().pmr_set_lu_cluster_panel(mode)
        Ad = nv.to_dev(A).clone()
        ipiv, info = eng.getrf(Ad)
        out.append((nv.to_host(Ad), nv.to_host(ipiv)[:777].copy(), int(nv.to_host(info)[0])))
    nv.lib().pmr_set_lu_cluster_panel(1)
    assert out[0][2] == 0 and out[1][2] == 0
    assert np.array_equal(out[0][1], out[1][1])
    assert np.abs(out[0][0] - out[1][0]).max() <= 1e-11


def test_getrf_cluster_panel_with_several_rows_per_thread():
    """n = 9000: the cluster-resident panels start with four register-resident rows per thread, then
    two, then one; the per-column path is the yardstick (same pivots, same factors)."""
    from pymolresponse_b200 import _engine as eng

    nv = _nv()
    torch = nv.torch_mod()
    n = 9000
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(n, n, dtype=torch.float64, device="cuda", generator=g)
    xt = torch.randn(n, 2, dtype=torch.float64, device="cuda", generator=g)
    b = A @ xt
    out = []
    for mode in (1, 0):
        nv.lib().pmr_set_lu_cluster_panel(mode)
        Ad = A.clone()
        ipiv, info = eng.getrf(Ad)
        assert int(info.item()) == 0
        out.append((Ad, ipiv.clone()))
    nv.lib().pmr_set_lu_cluster_panel(1)
    assert bool((out[0][1] == out[1][1]).all())
    assert float((out[0][0] - out[1][0]).abs().max()) <= 1e-9
    X = eng.getrs(out[0][0], out[0][1], b.clone())
    assert float((A @ X - b).abs().max()) <= 1e-7 * float(b.abs().max())


def test_getrf_flags_singular_matrix():
    from pymolresponse_b200 import _engine as eng

    nv = _nv()
    A = np.ones((6, 6))
    _, info = eng.getrf(nv.to_dev(A).clone())
    assert int(nv.to_host(info)[0]) != 0


# ----------------------------------------------------------------------------------- small helpers
def test_contract_and_energy_differences():
    from pymolresponse_b200 import utils

    r = _rng(18)
    P, R = r.standard_normal((3, 5000, 1)), r.standard_normal((4, 5000, 1))
    got = utils.form_results(P, R)
    assert got.shape == (3, 4)
    assert np.abs(got - P[:, :, 0] @ R[:, :, 0].T).max() <= 1e-11
    eo, ev = np.sort(r.uniform(-2, -1, 4)), np.sort(r.uniform(0.5, 2, 9))
    ed = utils.form_vec_energy_differences(eo, ev)
    want = np.array([ev[a] - eo[i] for i in range(4) for a in range(9)])
    assert np.array_equal(ed, want)


def test_jk_contract_matches_einsum():
    from pymolresponse_b200.integrals import JKDense

    r = _rng(19)
    n = 11
    I = r.standard_normal((n, n, n, n))
    D = r.standard_normal((5, n, n))
    jk = JKDense(I)
    for d in range(5):
        J, K = jk.compute_from_density(D[d])
        assert np.abs(J - np.einsum("pqrs,rs->pq", I, D[d])).max() <= 1e-12
        assert np.abs(K - np.einsum("prqs,rs->pq", I, D[d])).max() <= 1e-12
    Cl, Cr = r.standard_normal((n, 3)), r.standard_normal((n, 3))
    J, K = jk.compute_from_mocoeffs(Cl, Cr)
    Dm = Cl @ Cr.T
    assert np.abs(J - np.einsum("pqrs,rs->pq", I, Dm)).max() <= 1e-12
    assert np.abs(K - np.einsum("prqs,rs->pq", I, Dm)).max() <= 1e-12


# ----------------------------------------------------------------------------------- ERI upload
def _symmetric_eri(n, seed):
    T = _rng(seed).standard_normal((n, n, n, n))
    T = T + T.transpose(1, 0, 2, 3)
    T = T + T.transpose(0, 1, 3, 2)
    T = T + T.transpose(2, 3, 0, 1)
    return np.ascontiguousarray(T)


@pytest.mark.parametrize("layout", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 2, 7, 13, 32, 70])
def test_eri_upload_s8_rebuilds_the_full_tensor(n, layout):
    """Only the documented unique part is read from the host (everything else is poisoned with NaN
    here); the completed tensor in HBM equals the symmetric original bit for bit."""
    nv = _nv()
    T = _symmetric_eri(n, 41)
    keep = np.zeros((n, n, n, n), dtype=bool)
    if layout == 0:       # rows (mu, nu <= mu), columns up to (mu, mu)
        for mu in range(n):
            keep[mu, :mu + 1].reshape(mu + 1, n * n)[:, :mu * n + mu + 1] = True
    elif layout == 1:     # (mu >= la, nu | la, si <= la)
        for la in range(n):
            keep[la:, :, la, :la + 1] = True
    else:                 # rows (la, si <= la), columns from (la, 0) on
        for la in range(n):
            keep[la, :la + 1, la:, :] = True
    assert int(nv.lib().pmr_eri_s8_bytes(n, layout)) == 8 * int(keep.sum())
    H = np.where(keep, T, np.nan)
    got = nv.to_host(nv.upload_eri_s8(H, layout=layout))
    assert np.array_equal(got, T)


def test_ao2mo_s8_upload_matches_plain_upload():
    from pymolresponse_b200.ao2mo import AO2MO

    n, o = 12, 4
    T = _symmetric_eri(n, 42) / n
    Cm = np.linalg.qr(_rng(43).standard_normal((n, n)))[0]
    occ = (o, n - o, o, n - o)
    a = AO2MO(Cm, occ, I=T)
    a.perform_rhf_partial()
    b = AO2MO(Cm, occ, I=T, ao_symmetry="s8")
    b.perform_rhf_partial()
    for x, y in zip(a.tei_mo, b.tei_mo):
        assert np.array_equal(x, y)


def test_eri_slab_upload_ket_symmetry():
    """A nu-slab I[:, 2:5]: only the columns (la, si <= la) are read from the host."""
    nv = _nv()
    n = 9
    T = _symmetric_eri(n, 44)[:, 2:5].copy()
    keep = np.zeros(T.shape, dtype=bool)
    for la in range(n):
        keep[:, :, la, :la + 1] = True
    assert int(nv.lib().pmr_eri_ket_s2_bytes(T.shape[0] * T.shape[1], n)) == 8 * int(keep.sum())
    got = nv.to_host(nv.upload_eri_slab_s2(np.where(keep, T, np.nan)))
    assert np.array_equal(got, T)


@pytest.mark.parametrize("n,batch,nrhs", [(300, 1, 3), (700, 3, 4), (1537, 2, 1), (2560, 2, 6)])
def test_potrf_rows_fuses_the_forward_substitution(n, batch, nrhs):
    """pmr_potrf_rows: right-hand sides appended as extra ROWS leave the factorisation as
    (L^-1 b)^T; pmr_potrs_ex(phases=2) finishes the solve.  Against NumPy, look-ahead path included."""
    from pymolresponse_b200 import _engine as eng

    nv = _nv()
    r = _rng(n)
    A = _spd(n, n + 1, batch)
    Bm = r.standard_normal((batch, nrhs, n))
    aug = np.concatenate((A, Bm), axis=1)                       # (batch, n + nrhs, n)
    Sd = nv.to_dev(aug).clone()
    dinv, info = eng.potrf_rows(Sd, n, batch=batch)
    assert int(info.abs().sum().item()) == 0
    got = nv.to_host(Sd)
    for b in range(batch):
        Lref = np.linalg.cholesky(A[b])
        assert np.abs(np.tril(got[b, :n]) - Lref).max() <= 1e-11 * n
        Yref = np.linalg.solve(Lref, Bm[b].T)                   # (n, nrhs)
        assert np.abs(got[b, n:] - Yref.T).max() <= 1e-10 * n
    Y = Sd[:, n:, :].transpose(1, 2).contiguous()
    X = nv.to_host(eng.potrs_backward(Sd, n, dinv, Y, batch=batch))
    for b in range(batch):
        assert np.abs(A[b] @ X[b] - Bm[b].T).max() <= 1e-10 * n


def test_symmetrize_lower():
    from pymolresponse_b200 import _engine as eng

    nv = _nv()
    r = _rng(77)
    for n in (1, 31, 32, 33, 257):
        A = r.standard_normal((n, n))
        got = nv.to_host(eng.symmetrize_lower(nv.to_dev(A).clone()))
        want = np.tril(A) + np.tril(A, -1).T
        assert np.array_equal(got, want)
