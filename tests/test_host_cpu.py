"""CPU: the C-ABI library loads and exports every symbol include/pmr_b200.h declares; host-side
logic (shapes, index bookkeeping, DALTON labels, sharding helpers, 2-rank gloo exchange); the
product path refuses to compute without a GPU (no CPU fallback)."""

import json
import os
import re
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent


def _header_symbols():
    text = (REPO / "include" / "pmr_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pmr_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as entry

    entry.build()
    from pymolresponse_b200 import _native as nv

    lib = nv.load_library()
    names = _header_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in pmr_b200.h but not exported"
        assert name in nv.SYMBOLS, f"{name} has no ctypes signature"
    assert set(nv.SYMBOLS) == set(names)
    assert lib.pmr_version() >= 100
    assert lib.pmr_launch_count() == 0


def test_library_is_sm100a_with_dmma_and_async_copies():
    from pymolresponse_b200 import _native as nv

    out = subprocess.run(["cuobjdump", "-lelf", str(nv.library_path())], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    sass = subprocess.run(["cuobjdump", "-sass", str(nv.library_path())], capture_output=True,
                          text=True).stdout
    assert "DMMA.8x8x4" in sass and "LDGSTS" in sass


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pymolresponse_b200 import _native as nv
    from pymolresponse_b200 import explicit_equations_partial as eqp
    from pymolresponse_b200.ao2mo import AO2MO

    with pytest.raises(nv.PmrUnavailableError):
        eqp.form_rpa_b_matrix_mo_singlet_partial(np.zeros((2, 3, 2, 3)))
    a = AO2MO(np.eye(4), (1, 3, 1, 3), I=np.zeros((4, 4, 4, 4)))
    with pytest.raises(nv.PmrUnavailableError):
        a.perform_rhf_partial()


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|import_module\(.oracle|pmr_oracle", re.M)
    for path in (REPO / "pymolresponse_b200").rglob("*.py"):
        assert not pat.search(path.read_text()), path


def test_shape_helpers_and_index_maps():
    from pymolresponse_b200 import utils
    from pymolresponse_b200.indices import form_indices_from_occupations

    C = np.arange(12.0).reshape(3, 4)
    assert utils.fix_mocoeffs_shape(C).shape == (1, 3, 4)
    assert utils.fix_mocoeffs_shape((C, C)).shape == (2, 3, 4)
    e = np.array([1.0, 2.0, 3.0])
    E = utils.fix_moenergies_shape(e)
    assert E.shape == (1, 3, 3) and np.array_equal(np.diag(E[0]), e)
    assert utils.fix_moenergies_shape(np.stack((e, e))).shape == (2, 3, 3)
    assert utils.fix_moenergies_shape(np.diag(e)).shape == (1, 3, 3)
    m = np.arange(6.0).reshape(3, 2)   # (v, o): index i*v + a
    vec = utils.repack_matrix_to_vector(m)
    assert np.array_equal(vec, [0, 2, 4, 1, 3, 5])
    assert np.array_equal(utils.repack_vector_to_matrix(vec, (3, 2)), m)
    idx = form_indices_from_occupations((3, 4, 2, 5))
    assert idx.indices_closed_act == [(0, 2), (1, 2)]
    assert idx.indices_closed_secondary[0] == (0, 3) and len(idx.indices_closed_secondary) == 8
    assert idx.indices_act_secondary == [(2, 3), (2, 4), (2, 5), (2, 6)]


def test_enums_match_reference_names():
    from pymolresponse_b200.core import AO2MOTransformationType, Hamiltonian, Program, Spin

    assert Hamiltonian["RPA"].value == "rpa" and Hamiltonian["TDA"].value == "tda"
    assert Spin["singlet"].value == 0 and Spin["triplet"].value == 1
    assert AO2MOTransformationType.partial.value == "partial"
    assert {p.value for p in Program} >= {"pyscf", "psi4", "dalton"}


def test_dalton_labels():
    from pymolresponse_b200.interfaces.dalton.utils import clean_dalton_label, dalton_label_to_operator

    assert clean_dalton_label("PSO 002") == "pso_002"
    op = dalton_label_to_operator("XDIPLEN")
    assert (op.label, op.slice_idx, op.is_imaginary, op.is_spin_dependent) == ("dipole", 0, False, False)
    op = dalton_label_to_operator("ZANGMOM")
    assert (op.label, op.slice_idx, op.is_imaginary) == ("angmom", 2, True)
    op = dalton_label_to_operator("Y1SPNORB")
    assert (op.label, op.slice_idx, op.is_imaginary, op.is_spin_dependent) == ("spinorb1", 1, True, True)
    assert hasattr(op, "hsofac") and abs(op.hsofac - (0.0072973525693 ** 2) / 4) < 1e-12
    op = dalton_label_to_operator("FC H  02")
    assert (op.label, op.slice_idx, op.is_spin_dependent) == ("fermi", 1, True)
    op = dalton_label_to_operator("SD 004 y")
    assert (op.label, op.slice_idx) == ("sd", 6 + 1)


def test_fixture_readers(tmp_path):
    from pymolresponse_b200 import utils

    (tmp_path / "occ").write_text("2 4 2 4\n")
    assert utils.read_file_occupations(tmp_path / "occ") == (2, 4, 2, 4)
    (tmp_path / "m").write_text("2 3\n" + "\n".join(str(float(k)) for k in range(6)) + "\n")
    assert np.array_equal(utils.read_file_2(tmp_path / "m"), np.arange(6.0).reshape(2, 3))
    (tmp_path / "d").write_text("1 2 0.5\n2 2 1.5\n")
    assert np.array_equal(utils.parse_int_file_2(tmp_path / "d", 2), [[0, 0.5], [0.5, 1.5]])


def test_solver_constructor_contract():
    from pymolresponse_b200 import solvers
    from pymolresponse_b200.core import AO2MOTransformationType

    C, E = np.zeros((1, 4, 4)), np.zeros((1, 4, 4))
    s = solvers.ExactInv(C, E, (1, 3, 1, 3))
    assert not s.is_uhf and s.tei_mo is None and s.explicit_hessian is None
    assert s.tei_mo_type == AO2MOTransformationType.partial and s.inv_func is np.linalg.inv
    s.set_frequencies(None)
    assert s.frequencies == [0.0]
    with pytest.raises(AssertionError):
        solvers.ExactInv(np.zeros((4, 4)), E, (1, 3, 1, 3))
    with pytest.raises(AssertionError):
        solvers.ExactInv(np.zeros((2, 4, 4)), E, (1, 3, 1, 3))
    it = solvers.IterativeLinEqSolver(C, E, (1, 3, 1, 3), None)
    assert (it.maxiter, it.conv) == (40, 1.0e-9)


def test_shard_helpers():
    from pymolresponse_b200 import distributed as pd

    for n, w in ((512, 8), (10, 3), (3, 8)):
        parts = [pd.shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[k][1] == parts[k + 1][0] for k in range(w - 1))
        assert max(hi - lo for lo, hi in parts) - min(hi - lo for lo, hi in parts) <= 1
    assert pd.round_robin(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((pd.round_robin(10, r, 4) for r in range(4)), [])) == list(range(10))
    assert pd.rank_world() == (0, 1)


def test_flop_model_matches_survey():
    from pymolresponse_b200 import synthetic

    assert abs(synthetic.flops_ao2mo_rhf_partial(512, 64) / 1.460e13 - 1) < 2e-3
    nov = 64 * 448
    assert abs(synthetic.flops_solve_rhf(nov, 1, 7, 3) / 8.65e13 - 1) < 5e-3
    assert abs(synthetic.bytes_assemble(nov) / 26.3e9 - 1) < 2e-3


_WORKER = r"""
import sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[3])
from pymolresponse_b200 import distributed as pd
rank, world = int(sys.argv[1]), 2
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=rank, world_size=world)
assert pd.rank_world() == (rank, 2)
# partial MO-integral blocks are summed over ranks
blk = torch.full((3, 4), float(rank + 1), dtype=torch.float64)
pd.allreduce_blocks([blk])
assert torch.equal(blk, torch.full((3, 4), 3.0, dtype=torch.float64))
# frequencies dealt round-robin; every rank ends with all response vectors in frequency order
nf = 5
mine = pd.round_robin(nf, rank, world)
local = [torch.full((2, 6), float(f), dtype=torch.float64) for f in mine]
full = pd.gather_round_robin(local, nf)
assert len(full) == nf and all(torch.equal(full[f], torch.full((2, 6), float(f), dtype=torch.float64)) for f in range(nf))
# fewer items than ranks: rank 1 owns nothing
full = pd.gather_round_robin([torch.ones(2, 2, dtype=torch.float64)] if rank == 0 else [], 1)
assert len(full) == 1 and torch.equal(full[0], torch.ones(2, 2, dtype=torch.float64))
lo, hi = pd.shard_range(9, rank, world)
assert (lo, hi) == ((0, 5) if rank == 0 else (5, 9))
# a failure on one rank is seen by every rank before the next data-path collective
assert pd.any_rank(rank == 1) is True and pd.any_rank(False) is False
# row blocks of the MO integrals / Hessian: reduce-scatter of partial sums, all-gather of the blocks
# (even and ragged splits), column blocks of the shared K^-1
for rows in (6, 7):
    T = torch.arange(rows * 3, dtype=torch.float64).reshape(rows, 3) * (rank + 1)
    mine = pd.reduce_scatter_rows(T.clone(), rows)
    rlo, rhi = pd.shard_range(rows, rank, world)
    want = torch.arange(rows * 3, dtype=torch.float64).reshape(rows, 3) * 3.0
    assert torch.equal(mine, want[rlo:rhi])
    assert torch.equal(pd.allgather_rows(mine, rows), want)
cols = 8
clo, chi = pd.shard_range(cols, rank, world)
full = torch.arange(5 * cols, dtype=torch.float64).reshape(5, cols)
assert torch.equal(pd.allgather_columns(full[:, clo:chi].contiguous(), cols), full)
pd.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_exchange(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(port), str(REPO)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
             for r in range(2)]
    for r, p in enumerate(procs):
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert f"ok {r}" in out


def test_text_and_matrix_helpers(tmp_path):
    """The reference's tests/test_utils.py cases (Splitter) plus the small matrix helpers, checked
    against the reference's own functions when it is mounted."""
    from pymolresponse_b200 import utils

    sp = utils.Splitter((4, 3, 5, 6, 10, 10, 10, 10, 10, 10))
    full = "  60  H 10  s        0.14639   0.00000   0.00000  -0.00000  -0.00000   0.00000"
    short = "   1  C 1   s       -0.00000  -0.00000   0.00000"
    assert sp.split(full) == ["60", "H", "10", "s", "0.14639", "0.00000", "0.00000", "-0.00000",
                              "-0.00000", "0.00000"]
    assert sp.split(short) == ["1", "C", "1", "s", "-0.00000", "-0.00000", "0.00000"]
    assert sp.split(short, truncate=False) == ["1", "C", "1", "s", "-0.00000", "-0.00000", "0.00000",
                                               "", "", ""]
    ref = tmp_path / "ref.txt"
    ref.write_text("rpa singlet 0.0 xdiplen xdiplen 1.5\nrpa singlet 0.5 xdiplen ydiplen -2.25\n")
    assert utils.get_reference_value_from_file(ref, "rpa", "singlet", "0.5", "xdiplen", "ydiplen") == -2.25
    with pytest.raises(ValueError):
        utils.get_reference_value_from_file(ref, "tda", "singlet", "0.5", "xdiplen", "ydiplen")

    rng = np.random.default_rng(3)
    A = rng.standard_normal((5, 5))
    S, T = A + A.T, A - A.T
    assert (utils.matsym(S), utils.matsym(T), utils.matsym(A), utils.matsym(np.zeros((3, 3)))) == (1, 2, 0, 3)
    low, up = utils.flip_triangle_sign(A), utils.flip_triangle_sign(A, "upper")
    assert np.array_equal(np.triu(low, 1), np.triu(A, 1)) and np.array_equal(np.tril(low), -np.tril(A))
    assert np.array_equal(np.tril(up, -1), np.tril(A, -1)) and np.array_equal(np.triu(up), -np.triu(A))
    B = A.copy()
    B[0, 0] = 1e-20
    assert utils.screen(B)[0, 0] == 0.0 and B[0, 0] == 1e-20
    assert utils.form_indices_orbwin(2, 2) == [(0, 2), (0, 3), (1, 2), (1, 3)]
    assert utils.form_indices_zero(2, 2) == [(0, 0), (0, 1), (1, 0), (1, 1)]


def test_every_entry_point_is_documented():
    """INTEGRATION.md / DESIGN.md name every symbol include/pmr_b200.h declares."""
    hdr = (REPO / "include" / "pmr_b200.h").read_text()
    syms = set(re.findall(r"\b(pmr_[a-z0-9_]+)\s*\(", hdr))
    docs = (REPO / "INTEGRATION.md").read_text() + (REPO / "DESIGN.md").read_text()
    assert not [s for s in sorted(syms) if s not in docs]


def test_jacobi_tournament_schedule_covers_every_pair_once():
    """Host replica of pair_of() in csrc/pmr_eig.cu: Ne-1 steps of Ne/2 disjoint pairs visit every
    pair of columns exactly once per sweep (odd N: the padding player sits out)."""
    def pair_of(step, k, Ne):
        m = Ne - 1
        if k == 0:
            p, q = step % m, m
        else:
            p, q = (step + k) % m, (step - k % m + m) % m
        return (p, q) if p < q else (q, p)

    for N in (2, 3, 4, 7, 10, 33, 64):
        Ne = (N + 1) & ~1
        seen = set()
        for step in range(Ne - 1):
            used = set()
            for k in range(Ne // 2):
                p, q = pair_of(step, k, Ne)
                assert p != q and p not in used and q not in used   # disjoint within a step
                used.update((p, q))
                if q < N:
                    assert (p, q) not in seen
                    seen.add((p, q))
        assert len(seen) == N * (N - 1) // 2


def test_shared_inverse_block_partition():
    """Column blocks of the shared K^-1 (2*world blocks, rank r owns r and 2*world-1-r): a partition
    of [0, N), and the modelled substitution work (1-c)^2 + 1-c^2 per block is balanced."""
    from pymolresponse_b200 import distributed as pd

    for N, world in ((28672, 8), (7168, 4), (7168, 2), (1537, 2), (5, 4)):
        nblocks = 2 * world
        bounds = [pd.shard_range(N, b, nblocks) for b in range(nblocks)]
        assert bounds[0][0] == 0 and bounds[-1][1] == N
        assert all(bounds[i][1] == bounds[i + 1][0] for i in range(nblocks - 1))
        if N >= 1000:
            work = []
            for r in range(world):
                w_r = 0.0
                for lo, hi in (bounds[r], bounds[nblocks - 1 - r]):
                    c = lo / N
                    w_r += ((1 - c) ** 2 + 1 - c * c) * (hi - lo) / N
                work.append(w_r)
            assert max(work) / min(work) < 1.02


def test_eigensolver_error_behaviour_matches_reference():
    """UHF raises NotImplementedError (reference solvers.py:754-756, :854-856); running without MO
    integrals and without a program raises RuntimeError (solvers.py:781-784).  No GPU is touched."""
    from pymolresponse_b200 import solvers
    from pymolresponse_b200.core import AO2MOTransformationType, Hamiltonian, Spin

    n, na, nb = 6, 2, 1
    C = np.stack((np.eye(n), np.eye(n)))
    E = np.stack((np.diag(np.arange(n, dtype=float)), np.diag(np.arange(n, dtype=float))))
    for cls, ham in ((solvers.ExactDiagonalizationSolver, Hamiltonian.RPA),
                     (solvers.ExactDiagonalizationSolverTDA, Hamiltonian.TDA)):
        s = cls(C, E, (na, n - na, nb, n - nb))
        s.tei_mo = tuple(np.zeros((1, 1, 1, 1)) for _ in range(6))
        s.tei_mo_type = AO2MOTransformationType.partial
        with pytest.raises(NotImplementedError):
            s.form_explicit_hessian(ham, Spin.singlet, None)
        r = cls(C[:1], E[:1], (na, n - na, na, n - na))
        with pytest.raises(RuntimeError):
            r.run(ham, Spin.singlet, None, None)
    assert issubclass(solvers.ExactDiagonalizationSolverTDA, solvers.EigSolverTDA)
    assert issubclass(solvers.EigSolverTDA, solvers.EigSolver)
    x, y = solvers.EigSolver.norm_xy(np.array([3.0, 0.0, 0.0, 1.0]), 1, 2)
    assert abs(2 * (x @ x - y @ y) - 1.0) < 1e-15


def test_pyscf_and_psi4_adapters_with_duck_typed_packages(monkeypatch):
    """the integral adapters against stand-ins for pyscf.gto.Mole / psi4.core
    (neither package exists in this image): label -> intor name / component count, the spin-orbit
    sum over nuclei, occupations and the four gauge-origin keywords, psi4's MintsHelper hand-off."""
    from _fake_qc import FakeMole, install_fake_psi4

    from pymolresponse_b200 import integrals
    from pymolresponse_b200.interfaces import pyscf_adapter as pa

    n = 6
    mol = FakeMole(n, (3, 2))
    gen = pa.IntegralsPyscf(mol)
    assert np.array_equal(gen.integrals(integrals.DIPOLE), mol._ints["cint1e_r_sph"])
    assert mol.calls[-1] == ("cint1e_r_sph", 3, None)
    assert np.array_equal(gen.integrals(integrals.ANGMOM_COMMON_GAUGE), mol._ints["cint1e_cg_irxp_sph"])
    assert np.array_equal(gen.integrals(integrals.ANGMOM_GIAO), mol._ints["cint1e_giao_irjxp_sph"])
    assert np.array_equal(gen.integrals(integrals.DIPVEL), mol._ints["cint1e_ipovlp_sph"])
    so = gen.integrals(integrals.SO_SPHER_1e)
    want = sum(mol._charges[a] * mol._prinvxp[a] for a in range(3))
    np.testing.assert_allclose(so, want, rtol=0, atol=1e-14)
    assert np.abs(np.asarray(gen.integrals(integrals.SO_SPHER_1e_EFF))).max() == 0.0
    C = np.linalg.qr(np.random.default_rng(0).standard_normal((n, n)))[0][None]
    occ = pa.occupations_from_pyscf_mol(mol, C)
    assert occ == (3, 3, 2, 4)
    assert np.array_equal(pa.calculate_origin_pyscf("zero", mol, C, occ), np.zeros(3))
    com = pa.calculate_origin_pyscf("com", mol, C, occ)
    np.testing.assert_allclose(com, (mol._masses[:, None] * mol._coords).sum(0) / mol._masses.sum())
    ncc = pa.calculate_origin_pyscf("ncc", mol, C, occ)
    np.testing.assert_allclose(ncc, (mol._charges[:, None] * mol._coords).sum(0) / mol._charges.sum())
    ecc = pa.calculate_origin_pyscf("ecc", mol, C, (3, 3, 3, 3))
    D = 2 * C[0][:, :3] @ C[0][:, :3].T
    np.testing.assert_allclose(ecc, np.einsum("xpq,pq->x", mol._ints["cint1e_r_sph"], D) / np.trace(D))
    with pytest.raises(ValueError):
        pa.calculate_origin_pyscf("nonsense", mol, C, occ)
    # psi4
    dip, ang = install_fake_psi4(monkeypatch, n, eri=None)
    from pymolresponse_b200.interfaces.psi4_adapter import IntegralsPsi4

    g4 = IntegralsPsi4(molecule="h2o")
    assert np.array_equal(g4.integrals(integrals.DIPOLE), dip)
    assert np.array_equal(g4.integrals(integrals.ANGMOM_COMMON_GAUGE), ang)


def test_bench_helpers_without_gpu():
    """bench.py's host-side bookkeeping: workloads, the algorithmic flop model of SURVEY 8(d), the
    bounded reference sample, the config both arms print, the host-memory guard."""
    import bench

    w5 = bench.workload("config5", 1)
    assert (w5["n"], w5["o"], len(w5["pol_freqs"])) == (256, 32, 64) and w5["magnetizability"]
    # config 5 F_total = 8.68e12 (+ the static K-solve of the imaginary operator)
    assert abs(bench.algorithmic_flops(w5) - 8.72e12) < 0.02e12
    w3 = bench.workload("config3", 8)
    assert abs(bench.algorithmic_flops(w3) - 1.01e14) < 0.01e14
    w4 = bench.workload("config4", 2)
    assert w4["input"] == "dense" and abs(bench.algorithmic_flops(w4) - 1.84e13) < 0.03e13
    assert bench.workload("config4", 1)["input"] == "df"
    # both arms print the same static config (the driver compares them)
    assert bench.static_config(w5, 1) == bench.static_config(bench.workload("config5", 1), 1)
    assert "fp64" not in json.dumps(bench.static_config(w5, 1))     # no measured values inside
    # the reference sample grows with the per-step budget
    assert bench.cpu_sample_spec(20, 5)["n"] == 128
    assert bench.cpu_sample_spec(3, 3)["n"] == 160
    assert bench.cpu_sample_spec(1000, 1000)["n"] == 64
    a = bench.host_memory_available()
    assert a is None or a > 0


def test_factorised_sample_indices_are_shared_between_bench_and_oracle():
    from oracle import pmr_factorised as fac

    r1, c1 = fac.sample_indices(7168, 64, 7)
    r2, c2 = fac.sample_indices(7168, 64, 7)
    assert np.array_equal(r1, r2) and np.array_equal(c1, c2)
    assert (r1 >= c1).all() and r1.max() < 7168 and len(r1) == 64 + 64
