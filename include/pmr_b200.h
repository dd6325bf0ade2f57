/*
 * pmr_b200.h -- C ABI of libpmr_b200.so, the B200 (sm_100a) implementation of
 * pymolresponse's linear-response hot path.
 *
 * The reference (berquist/pymolresponse) is pure Python/NumPy: there is no FFI in it.  The
 * seam this library sits behind is the reference's Python function/class API; each entry
 * point below names the reference call site (file:line under /root/reference/pymolresponse/)
 * whose arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to FP64 data unless the name ends in _host;
 *   - all matrices are C-order (row-major) with explicit element strides;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it;
 *   - return value: 0 on success, negative pmr_status on error (pmr_last_error() has text);
 *   - factorisation failure (non-SPD / singular) is reported through a device int `info`
 *     (0 = ok, k>0 = first bad pivot, 1-based), never through the return value.
 *   - no global state except a per-thread error string; thread-safe per stream.
 */
#ifndef PMR_B200_H
#define PMR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PMR_OK = 0,
  PMR_ERR_ARG = -1,      /* bad argument (shape / stride / null pointer) */
  PMR_ERR_CUDA = -2,     /* CUDA runtime error (launch, attribute, ...)   */
  PMR_ERR_WORKSPACE = -3,/* workspace too small                           */
  PMR_ERR_CONVERGENCE = -4 /* an iterative kernel driver hit its sweep limit */
} pmr_status;

int pmr_version(void);
const char* pmr_last_error(void);
/* Number of kernels this library launched from the calling thread since the last reset. */
long long pmr_launch_count(void);
void pmr_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * FP64 DMMA GEMM:  C(m,n) = alpha * sum_k A(m,k) B(k,n) + beta * C(m,n), batched two levels.
 * A(m,k) = A[m*a_sm + k*a_sk], B(k,n) = B[k*b_sk + n*b_sn], C(m,n) = C[m*c_sm + n*c_sn];
 * for A and for B one of the two strides must be 1.  This is the building block behind
 *   ao2mo.py:45-49 (np.dot + three einsum quarter transforms),
 *   solvers.py:505,527-530 (inverse / Cholesky trailing updates),
 *   operators.py:85-87 (C_v^T O C_o), utils.py:26 (P R^T).
 * ---------------------------------------------------------------------------------------- */
#define PMR_GEMM_LOWER 1      /* write only C(m,n) with n <= m; tiles above the diagonal skipped */
#define PMR_GEMM_KSTART_M 2   /* operands vanish for k < m-tile origin: start the k loop there  */
#define PMR_GEMM_KEND_M 4     /* operands vanish for k >= m-tile end: stop the k loop there     */
#define PMR_GEMM_C_ALIASES_A 8 /* C is A (in-place A <- alpha A B, beta = 0); needs N <= 128       */

typedef struct {
  int M, N, K;
  double alpha, beta;
  const double* A;
  long long a_sm, a_sk;
  const double* B;
  long long b_sk, b_sn;
  double* C;
  long long c_sm, c_sn;
  int batch1, batch2;                           /* >= 1 each */
  long long a_b1, a_b2, b_b1, b_b2, c_b1, c_b2; /* element strides per batch index */
  int flags;
  /* optional second copy of the result with its own strides (e.g. transposed); NULL = none.
   * Honoured by the cp.async kernels (always used with PMR_GEMM_C_ALIASES_A). */
  double* C2;
  long long c2_sm, c2_sn, c2_b1, c2_b2;
} pmr_gemm_t;

int pmr_dgemm(const pmr_gemm_t* g, void* stream);
/* Measurement aid (bench.py roofline pass): while enabled every GEMM launch of the calling thread
 * is bracketed by CUDA events on its own stream; fetch sums them per tile class
 * (out_host[cls*4 + {0,1,2,3}] = launches, ms, algorithmic flops, algorithmic bytes).
 * Disabling frees the events.  Off by default; never enabled inside a timed region. */
void pmr_profile_enable(int on);
int pmr_profile_fetch(double* out_host, int ncls);

/* ------------------------------------------------------------------------------------------
 * AO -> MO two-electron integral transformation.
 * pmr_ao2mo_transform: AO2MO.transform (ao2mo.py:35-50), quarters in index order 1,2,3,4.
 *   I is (n,n,n,n); Ck is (n x dk) with row stride ldck; out is (d1,d2,d3,d4) C-order.
 * pmr_ao2mo_partial: the occupied-first chain that produces the blocks of
 *   AO2MO.perform_rhf_partial (ao2mo.py:58-70) / AO2MOpyscf.perform_uhf_partial
 *   (interfaces/pyscf/ao2mo.py:71-109) from ONE shared first quarter per bra spin:
 *     ovov_k[i,a,j,b] = (i a | j_k b_k)  for k < n_ket   (ket spin k has its own Co2/Cv2)
 *     oovv  [i,j,a,b] = (i j | a b)      if oovv != NULL (bra spin only)
 *   I_slab holds I[:, nu_lo:nu_hi, :, :] as (n, nu_cnt, n, n) C-order: the AO ERI tensor sharded
 *   over the second index of the first AO pair; partial sums over the local nu are ACCUMULATED
 *   into ovov/oovv when `accumulate` != 0 (ranks then sum with one reduce per block).
 * Workspace sizes in doubles come from the *_workspace functions.
 * ---------------------------------------------------------------------------------------- */
long long pmr_ao2mo_transform_workspace(int n, int d1, int d2, int d3, int d4);
int pmr_ao2mo_transform(const double* I, int n, const double* C1, int ldc1, int d1,
                        const double* C2, int ldc2, int d2, const double* C3, int ldc3, int d3,
                        const double* C4, int ldc4, int d4, double* out, double* workspace,
                        long long workspace_doubles, void* stream);

typedef struct {
  const double* Co2; /* (n x o2) ket occupied coefficients, row stride ldc2 */
  const double* Cv2; /* (n x v2) ket virtual coefficients, row stride ldc2  */
  int ldc2, o2, v2;
  double* ovov; /* (o1, v1, o2, v2) C-order output */
} pmr_ao2mo_ket_t;

long long pmr_ao2mo_partial_workspace(int n, int nu_cnt, int o1, int v1, int n_ket,
                                      const pmr_ao2mo_ket_t* kets, int want_oovv, int nu_block);
/* accumulate: bit 0 = add into ovov/oovv; bit 1 = stop the (ij|ab) chain after its two occupied
 * quarters and return the half-transformed (i j|la si) in `oovv` (then o1*o1*n*n doubles): ranks
 * reduce-scatter that over i and finish only their own rows with pmr_ao2mo_oovv_tail. */
int pmr_ao2mo_partial(const double* I_slab, int n, int nu_lo, int nu_cnt, const double* Co1,
                      const double* Cv1, int ldc1, int o1, int v1, int n_ket,
                      const pmr_ao2mo_ket_t* kets, double* oovv, int accumulate, int nu_block,
                      double* workspace, long long workspace_doubles, void* stream);

/* Last two (virtual) quarters of (ij|ab) for `rows` leading occupied indices:
 * Y is (rows, o1, n, n), oovv is (rows, o1, v1, v1); workspace from the _workspace function. */
long long pmr_ao2mo_oovv_tail_workspace(int n, int rows, int o1, int v1);
int pmr_ao2mo_oovv_tail(const double* Y, int n, const double* Cv1, int ldc1, int rows, int o1,
                        int v1, double* oovv, int accumulate, double* workspace,
                        long long workspace_doubles, void* stream);

/* ------------------------------------------------------------------------------------------
 * Orbital-Hessian block assembly (explicit_equations_partial.py:8-214,
 * explicit_equations_full.py:8-238).  One bandwidth-bound gather kernel:
 *   P[ia,jb] = c1 * T1[(i,a,j,b)] - c2 * T2[(i,a,j,b)] + delta(ia,jb) * (ev[a] - eo[i])
 *   Q[ia,jb] = -(c1q * U1[(i,a,j,b)] - c2q * U2[(i,a,j,b)])
 * with each operand addressed by four element strides (partial AND full layouts are both
 * expressible).  Outputs: out_mode 0 -> (A = P, B = Q) ; out_mode 1 -> (M = P - Q, K = P + Q),
 * the A-B / A+B matrices the half-size solver consumes.  Either output pointer may be NULL.
 * ia = i*nv_x + a (a fast), jb = j*nv_y + b.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const double* ptr; /* NULL => term absent */
  long long s_i, s_a, s_j, s_b;
  double coef;
} pmr_gather_term_t;

typedef struct {
  int no_x, nv_x, no_y, nv_y; /* row space (i,a), column space (j,b) */
  pmr_gather_term_t p1, p2;   /* P = p1.coef*p1 - p2.coef*p2 (+ diagonal) */
  pmr_gather_term_t q1, q2;   /* Q = -(q1.coef*q1 - q2.coef*q2) */
  const double* eps_occ;      /* NULL => no diagonal */
  const double* eps_virt;
  int out_mode;               /* 0: (P,Q)  1: (P-Q, P+Q) */
  int i_lo, i_hi;             /* only rows with i in [i_lo, i_hi) are produced (row-block shard) */
  double* out0;
  double* out1;
  long long ld0, ld1;         /* row strides of the outputs (row index is global ia) */
} pmr_assemble_t;

int pmr_assemble(const pmr_assemble_t* a, void* stream);
/* ediff[i*nv + a] = eps_virt[a] - eps_occ[i]  (utils.py:326-347) */
int pmr_energy_differences(const double* eps_occ, int no, const double* eps_virt, int nv,
                           double* ediff, void* stream);

/* ------------------------------------------------------------------------------------------
 * Dense factorisations and solves (replacing np.linalg.inv / scipy cholesky+inv,
 * solvers.py:499-555, and the G^-1 . rhs products, solvers.py:401-466).
 * All matrices row-major; `batch` independent problems at element stride `stride_*`.
 * Cholesky: lower factor L (A = L L^T) overwrites the lower triangle; the strictly upper
 * triangle is neither read nor written.  `dinv` receives the inverses of the 128x128
 * diagonal blocks of L (needed by the solves): pmr_potrf_dinv_doubles(N) doubles per matrix.
 * ---------------------------------------------------------------------------------------- */
long long pmr_potrf_dinv_doubles(int N);
/* pmr_potrf overlaps panel factorisations with trailing updates on a second, high-priority stream
 * (look-ahead).  Switch it off to get serialised, individually timeable kernels (bench.py's
 * roofline pass does); results are identical either way.  Per calling thread; default on. */
void pmr_set_lookahead(int on);
/* Development aid: while a device buffer (>= 64 long long) is set, the diagonal-block kernel of the
 * Cholesky (tools/potf2_phases.py) and the first cluster-resident inner panel of pmr_getrf
 * (tools/lu_panel_phases.py) store clock64() at their phase boundaries; NULL switches it off. */
int pmr_debug_potf2_timing(long long* device_buffer);
int pmr_potrf(double* A, int N, long long lda, int batch, long long stride_a, double* dinv,
              int* info /* device, batch ints */, void* stream);
/* Same factorisation of the leading N x N block of a (rows x N) matrix, rows >= N: the rows below
 * the square ride through every panel solve and trailing update, so right-hand sides stored there
 * (one per ROW) leave as Y^T with L Y = B -- the forward substitution of G^-1 . rhs
 * (solvers.py:407-410) costs no extra pass over the factor; finish with pmr_potrs_ex(phases = 2). */
int pmr_potrf_rows(double* A, int N, int rows, long long lda, int batch, long long stride_a,
                   double* dinv, int* info, void* stream);
/* pmr_potrs with the two substitutions selectable: phases 1 = forward (B -> work, work holds Y as
 * (N, nrhs) per matrix), 2 = backward (work -> B), 3 = both. */
int pmr_potrs_ex(const double* L, int N, long long lda, int batch, long long stride_l,
                 const double* dinv, double* B, int nrhs, long long ldb, long long stride_b,
                 double* work, int phases, void* stream);
/* A[c][r] = A[r][c] for r > c: completes a symmetric matrix of which only the lower triangle was
 * computed (K^-1 from pmr_potri_lower, then used as a plain GEMM operand: p = w K^-1 m). */
int pmr_symmetrize_lower(double* A, int N, long long lda, void* stream);
/* Factor ONE 512-wide block column (columns [j0, j0+512), j0 % 512 == 0) of a single matrix whose
 * rows j0.. already carry the updates of the block columns to its left; `dinv` / `info` as in
 * pmr_potrf (info must be zeroed by the caller once).  Multi-GPU building block. */
int pmr_potrf_panel(double* A, int N, long long lda, int j0, double* dinv, int* info, void* stream);
/* Solve A X = B in place of B (B is N x nrhs, row stride ldb); `work` has N*nrhs doubles/batch. */
int pmr_potrs(const double* L, int N, long long lda, int batch, long long stride_l,
              const double* dinv, double* B, int nrhs, long long ldb, long long stride_b,
              int first_row /* rows of B above it are zero on entry (0 = general) */,
              int lower_only /* 1: rows of X above first_row are not computed (left zero) */,
              double* work, void* stream);
/* 1 (default): the four dependent 128-row steps inside every 512-row chunk of a substitution run
 * in ONE kernel (running right-hand side in shared memory) instead of eight small GEMM launches. */
void pmr_set_fused_substitution(int on);
/* The two substitutions of pmr_potrs on their own (one matrix): the forward pass of the 512-row
 * chunks [chunk_lo, chunk_hi) only needs the factor's columns up to chunk_hi*512, so it can follow a
 * factorisation that is still running (the distributed Cholesky hands over one block column at a
 * time); `work` (N*nrhs doubles, zero on first use) carries Y to the backward call. */
int pmr_potrs_phase(const double* L, int N, long long lda, const double* dinv, double* B, int nrhs,
                    long long ldb, int first_row, int lower_only, double* work,
                    int backward /* 0: forward chunks, 1: backward pass */, int chunk_lo,
                    int chunk_hi, void* stream);
/* Lower triangle of A^-1 from the Cholesky factor (out may not alias L); work: N*N doubles. */
int pmr_potri_lower(const double* L, int N, long long lda, const double* dinv, double* out,
                    long long ldo, double* work, void* stream);
/* Z = inv(L) for the lower Cholesky factor (diagonal-block inverses in `dinv`) by recursive
 * doubling: log2(N/128) levels of two batched GEMMs (Z21 = -Z22 L21 Z11 for every pair of
 * neighbouring blocks) instead of one substitution step per 128 columns.  Entries of Z above its
 * 128-wide block diagonal are scratch on exit.  pmr_gram_lower_cols: columns [c0, c0+ncols) of the
 * lower triangle of Z^T Z (= (L L^T)^-1; c0 % 128 == 0) -- together the scipy.linalg.inv(chol)
 * + product of ExactInvCholesky (solvers.py:527-530), and the column blocks ranks share out. */
int pmr_trtri_tree(const double* L, int N, long long lda, const double* dinv, double* Z,
                   long long ldz, void* stream);
int pmr_gram_lower_cols(const double* Z, int N, long long ldz, int c0, int ncols, double* out,
                        long long ldo, void* stream);
/* S_f = M + shift_a[f] * I + shift_b[f] * W on the lower triangle, f < nf (frequency batch). */
int pmr_shifted_combine(const double* M, const double* W, int N, long long ldm, long long ldw,
                        const double* shift_a_host, const double* shift_b_host, int nf,
                        double* S, long long lds, long long stride_s, void* stream);
/* A[i][i] += value, i < N (frequency shift of the explicit super-Hessian, solvers.py:230-231) */
int pmr_shift_diagonal(double* A, int N, long long lda, double value, void* stream);
/* LU with partial pivoting (general / above-pole case): row-major A = P L U in place, LAPACK's pivots.
 * Two levels of blocking: 64-column panels are applied inside a 256-column block, the trailing matrix
 * gets one interchange pass (as a single row permutation), one U12 solve and one rank-256 update per
 * block. */
int pmr_getrf(double* A, int N, long long lda, int* ipiv /* device, N */, int* info,
              double* work /* pmr_getrf_work_doubles(N) */, void* stream);
long long pmr_getrf_work_doubles(int N);
/* Panels of the pivoted LU are factorised eight columns per launch by one thread-block cluster (rows in
 * registers, one cluster barrier per column) while N - k0 <= 16384 rows remain, and one launch per
 * column otherwise.  Switching the cluster panel off forces the per-column path (the tests run both;
 * the pivots and the factors are the same).  Default on. */
void pmr_set_lu_cluster_panel(int on);
int pmr_getrs(const double* LU, int N, long long lda, const int* ipiv, double* B, int nrhs,
              long long ldb, double* work /* N*nrhs + N doubles */, void* stream);

/* ------------------------------------------------------------------------------------------
 * Operator right-hand sides and property contraction
 * (operators.py:57-125, utils.py:17-27, cphf.py:68-253).
 * ---------------------------------------------------------------------------------------- */
/* rhs[c, 0:nov] = scale * vec(C_v^T O_c C_o) (index i*nv+a), rhs[c, nov:2nov] = sign * same.
 * mask (device, nov ints or NULL): entries with mask != 0 are zeroed (triplet operators). */
int pmr_form_rhs(const double* O, int ncomp, int n, const double* C, int ldc, int nocc,
                 double sign, double scale, const int* mask, double* rhs /* ncomp x 2nov */,
                 double* rhs_ai /* ncomp x nov or NULL */, double* workspace /* ncomp*n*nocc */,
                 void* stream);
/* out[a,b] = pref * sum_k P[a,k] R[b,k]; partial: nblocks*nP*nR doubles of workspace. */
long long pmr_contract_workspace(int nP, int nR, long long len);
int pmr_contract_results(const double* P, int nP, long long ldp, const double* R, int nR,
                         long long ldr, long long len, double pref, double* out /* nP x nR */,
                         double* workspace, void* stream);
/* Elementwise helpers used by the half-size solver:
 *   xy[c,0:nov] = (p+m)/2, xy[c,nov:2nov] = (p-m)/2   (supervector [X;Y], solvers.py:418-425) */
int pmr_pm_to_xy(const double* p, const double* m, int ncomp, long long nov, long long ldp,
                 long long ldm, double* xy, long long ldxy, void* stream);
/* Packing kernels of the half-size solver (right-hand sides [g; s g] of solvers.py:401-425 for all
 * frequencies and operator components at once; host arrays: imag_col[c] = -1 for a real operator,
 * else the column of T = K^-1 g; w = the frequencies of this call, <= 64 each):
 *   pmr_solve_pack_rhs  out[q][c][k] = 2 g[c][k] | 2 w_q T[k][imag_col[c]] | 0 (padding), written as
 *                       the extra ROWS (row stride ldo) of the matrices pmr_potrf_rows factorises
 *                       (stride_q apart)
 *   pmr_solve_pack_p    prhs[k][q*ncols + c] = w_q m[q][k][c] + (imaginary ? 2 g[c][k] : 0)
 *   pmr_solve_xy        xy[(q,c)] = [(p + m)/2 ; (p - m)/2], p[k][q*ncols + c] (or NULL), m[q][k][c] */
int pmr_solve_pack_rhs(const double* g, long long ldg, const double* T, int nimag,
                       const int* imag_col_host, int ncols, int ncp, const double* w_host, int nfc,
                       long long N, double* out, long long ldo, long long stride_q, void* stream);
int pmr_solve_pack_p(const double* m, int ncp, const double* g, long long ldg,
                     const int* imag_col_host, int ncols, const double* w_host, int nfc, long long N,
                     double* prhs, void* stream);
int pmr_solve_xy(const double* p, const double* m, int ncp, int ncols, int nfc, long long N,
                 double* xy, void* stream);
/* y = a*x + b*y on ncols x nrows strided blocks */
int pmr_axpby(long long rows, long long cols, double a, const double* x, long long ldx, double b,
              double* y, long long ldy, void* stream);
/* out[c,k] = in[c,k] / (ediff[k mod nov] -/+ w)  (uncoupled response, cphf.py:94-141) */
int pmr_uncoupled_divide(const double* sv, int ncomp, long long nov, const double* ediff, double w,
                         double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Matrix-free path (IterativeLinEqSolver, solvers.py:603-630; JK contract integrals.py:41-67):
 *   J_d[p,q] = sum_rs I[p,q,r,s] D_d[r,s],  K_d[p,q] = sum_rs I[p,r,q,s] D_d[r,s]
 * for nd densities in ONE pass over the AO ERI tensor.
 * ---------------------------------------------------------------------------------------- */
int pmr_jk_contract(const double* I, int n, const double* D /* nd x n x n */, int nd,
                    double* J /* nd x n x n */, double* K /* nd x n x n */, void* stream);

/* One fixed-point sweep of the reference's iterative CPHF (solvers.py:621-637) for ncomp
 * (n x n) amplitude matrices x: on the occupied-virtual and virtual-occupied blocks
 *   x += (rhs - (x * d - G)) / d,   d[p,q] = eps[p] - eps[q] - omega,  G = C^T (2J1-K1+2J2-K2) C
 * and maxdiff[c] = max over all elements of (x_new - x_old) (signed, as the reference). */
int pmr_cphf_update(double* x, const double* x_old, const double* rhs, const double* G,
                    const double* eps, double omega, int n, int nocc, int ncomp,
                    int init /* 1: x = uncoupled guess rhs/(+-(e_a-e_i-w)), solvers.py:606-611 */,
                    double* maxdiff /* device, ncomp */, void* stream);

/* ------------------------------------------------------------------------------------------
 * Matrix-free block PCG for the iterative path (replaces the fixed-point loop of
 * IterativeLinEqSolver._run, solvers.py:603-637, when method="cg"): solves
 *   [[K, -w], [-w, M]] [p; m] = [(1+s) g; (1-s) g],  K = A + B_ref, M = A - B_ref
 * for k systems at once without forming A or B: sigma = H d comes from three pmr_dgemm calls on
 * the MO blocks (U = (ia|jb) d, W = (ib|ja) d, V = (ij|ab) d; explicit_equations_partial.py:34-38,
 * 92-96) and the kernels below.  Vectors are (nov, 2k) row-major, columns [0,k) = p parts,
 * [k,2k) = m parts; `w` is a device array of k frequencies; all CG scalars are device arrays.
 *   pmr_iter_permute_ov_rows   Zp[(b,j),:] = Z[(j,b),:]            (operand order of the W GEMM)
 *   pmr_iter_sigma_combine     rows [row0,row0+rows) of sigma from local U/V/W rows and global d:
 *                              q_p = cK U_p - V_p + sK W_p + D d_p - w d_m,
 *                              q_m = cM U_m - V_m + sM W_m + D d_m - w d_p   (D = e_a - e_i)
 *   pmr_iter_precond           y = P^-1 r, P = [[D,-w],[-w,D]] (the uncoupled Hessian,
 *                              cphf.py:94-96: the reference's initial guess, solvers.py:606-611)
 *   pmr_iter_coldot            out[c] = sum_rows X[:,c] Y[:,c] + X[:,k+c] Y[:,k+c] (deterministic)
 *   pmr_iter_update_zr         alpha = rho/dq;  z += alpha d;  r -= alpha q
 *   pmr_iter_update_d          beta = rho_new/rho_old;  d = y + beta d
 * ---------------------------------------------------------------------------------------- */
int pmr_iter_permute_ov_rows(const double* Z, int no, int nv, int ncols, double* Zp, void* stream);
int pmr_iter_sigma_combine(const double* U, const double* V, const double* W, const double* Z,
                           const double* ediff, const double* w, long long row0, long long rows,
                           int k, double cK, double sK, double cM, double sM, double* q,
                           void* stream);
int pmr_iter_precond(const double* r, const double* ediff, const double* w, long long nov, int k,
                     double* y, void* stream);
long long pmr_iter_coldot_workspace(long long nov, int k);
int pmr_iter_coldot(const double* X, const double* Y, long long nov, int k, double* out,
                    double* work, void* stream);
int pmr_iter_update_zr(double* z, double* r, const double* d, const double* q, const double* rho,
                       const double* dq, long long nov, int k, void* stream);
int pmr_iter_update_d(double* d, const double* y, const double* rho_new, const double* rho_old,
                      long long nov, int k, void* stream);

/* ------------------------------------------------------------------------------------------
 * AO ERI upload exploiting the 8-fold permutational symmetry (host buffer in, full tensor in HBM).
 * Replaces the plain host->device copy of the (n,n,n,n) tensor the reference hands to
 * AO2MO.__init__ (pymolresponse/ao2mo.py:18-33; the tensors come from pyscf int2e / psi4 ao_eri,
 * solvers.py:86-127, and always carry this symmetry).
 * layout 0 ("row prefix", ~n^4/3 doubles): for each mu the rows (mu, 0..mu), columns up to
 *   (mu, mu) -- long contiguous runs;  layout 1 ("box", ~n^4/6 doubles): for each la the entries
 *   (mu >= la, nu | la, si <= la) -- half the bytes in runs of (la+1) doubles;  layout 2 ("tail",
 *   the default): the pair-swapped image of the box, rows (la, si <= la), columns from (la, 0) to
 *   the end -- the same sixth of the doubles in long runs.
 *   pmr_eri_s8_bytes     bytes that cross PCIe for basis size n
 *   pmr_eri_upload_s8    one strided DMA per mu (layout 0) / per la (layouts 1, 2);
 *                        host: full C-order (n,n,n,n) array (pinned for an asynchronous copy)
 *   pmr_eri_complete_s8  fills the rest of the tensor in place from the uploaded part
 * ------------------------------------------------------------------------------------------ */
long long pmr_eri_s8_bytes(int n, int layout);
int pmr_eri_upload_s8(const double* host, double* dev, int n, int layout, void* stream);
int pmr_eri_complete_s8(double* dev, int n, int layout, void* stream);
/* Slab of `rows` (mu, nu) pairs, each an n x n matrix in (la, si) -- e.g. one rank's nu-shard
 * I[:, nu_lo:nu_hi] of a sharded tensor: only (la si| = (si la| is usable inside a slab; the
 * columns (la, si <= la) are copied (half the bytes) and every row is symmetrised in HBM. */
long long pmr_eri_ket_s2_bytes(long long rows, int n);
int pmr_eri_upload_ket_s2(const double* host, double* dev, long long rows, int n, void* stream);
int pmr_eri_complete_ket_s2(double* dev, long long rows, int n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Excitation energies: one-sided Jacobi SVD  G V = U Sigma  (replaces scipy.linalg.eig on the
 * explicit RPA / TDA Hessian, reference solvers.py:758-776 and :841-856; the host side reduces
 * the RPA problem to G = chol(A-B)^T chol(A+B), see pymolresponse_b200/solvers.py).
 *   Gt, Vt   (N, N) row-major, row p = COLUMN p of G / V; Vt = identity on entry.  On exit the
 *            rows of Gt are sigma_p u_p and the rows of Vt the right singular vectors (unsorted).
 *   sigma    device, N: singular values;  gv: device, N: g_p . v_p (= signed eigenvalue when G is
 *            symmetric);  scratch: device, >= 1 double.
 *   The call synchronises the stream once per sweep to read the convergence scalar
 *   max |g_p . g_q| / (|g_p| |g_q|); returns PMR_ERR_CONVERGENCE past max_sweeps.
 * pmr_zero_upper: A <- tril(A), so that a Cholesky factor can be a plain GEMM operand.
 * ------------------------------------------------------------------------------------------ */
int pmr_jacobi_svd(double* Gt, double* Vt, int N, long long ld, int max_sweeps, double tol,
                   double* sigma, double* gv, double* scratch, int* sweeps_done /* host */,
                   double* final_off /* host */, void* stream);
int pmr_zero_upper(double* A, int N, long long lda, void* stream);
/* column scalings of a row-major (rows, cols) matrix: A[:, j] *= scale[j]; unit 2-norm columns
 * with the component of largest magnitude made positive (LAPACK-style eigenvector columns). */
int pmr_scale_columns(double* A, int rows, int cols, long long lda, const double* scale,
                      void* stream);
int pmr_normalize_columns(double* A, int rows, int cols, long long lda, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PMR_B200_H */
