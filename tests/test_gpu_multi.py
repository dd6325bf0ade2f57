"""GPU: the N > 1 path with two ranks.  Both ranks use cuda:0 and exchange through gloo (host
staged), so it runs on the single-GPU box too; on multi-GPU boxes bench.py exercises NCCL.
Covers: nu-sharded AO2MO + block all-reduce, frequencies dealt to ranks, the shared K^-1
(column blocks solved per rank + all-gather), a rank without any frequency, gather of response
vectors -- compared with the oracle."""

import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent

_WORKER = r"""
import sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[3])
from oracle import pmr_oracle as orc
from pymolresponse_b200 import _engine as eng, _native as nv
from pymolresponse_b200 import cphf, distributed as pd, operators, solvers, synthetic
from pymolresponse_b200.ao2mo import AO2MO
from pymolresponse_b200.core import Hamiltonian, Spin

rank, world = int(sys.argv[1]), 2
torch.cuda.set_device(0)
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=rank, world_size=world)
# block-cyclic distributed Cholesky (panels broadcast) against the single-rank factorisation
for N in (700, 1537):
    r = np.random.default_rng(5)
    X = r.standard_normal((N, N))
    A = X @ X.T / N + 1.5 * np.eye(N)
    Ad = nv.to_dev(A).clone()
    dinv, info = eng.potrf_distributed(Ad)
    assert int(info.item()) == 0
    Lm = np.tril(Ad.cpu().numpy())
    assert np.abs(Lm @ Lm.T - A).max() <= 1e-12 * N
    rhs = r.standard_normal((N, 3))
    Xs = eng.potrs(Ad, dinv, nv.to_dev(rhs).clone()).cpu().numpy()
    assert np.abs(A @ Xs - rhs).max() <= 1e-10 * N
A[900, 900] = -3.0
_, info = eng.potrf_distributed(nv.to_dev(A).clone())
assert int(info.item()) == 901
eng.DISTRIBUTED_POTRF_MIN_N = 0      # let the small solver case below take the distributed path too
eng.KINV_TREE_MAX_N = 0              # (distributed factor of K + substitution-based shared K^-1)
n, o = 30, 4
case = synthetic.rhf_case(n, o)
lo, hi = pd.shard_range(n, rank, world)
a = AO2MO(case["C"], case["occupations"], I=np.ascontiguousarray(case["I"][:, lo:hi]),
          nu_range=(lo, hi), device_results=True, reduce=True)
a.perform_rhf_partial()
tei_ref = orc.ao2mo_rhf_partial(case["I"], case["C"], o)
for got, ref in zip(a.tei_mo, tei_ref):
    assert np.abs(got.cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max()
Or = synthetic.operator_integrals(n, 3, False)
Oi = synthetic.operator_integrals(n, 2, True)
for freqs in ([0.0, 0.1, 0.2], [0.3], [0.0]):
    solver = solvers.ExactInv(case["C"], case["E"], case["occupations"])
    solver.distribute_frequencies = True
    solver.tei_mo = a.tei_mo
    driver = cphf.CPHF(solver)
    ops = [operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or),
           operators.Operator(label="angmom", is_imaginary=True, ao_integrals=Oi)]
    for op in ops:
        driver.add_operator(op)
    driver.set_frequencies(freqs)
    driver.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    ref = orc.run_cphf(case["C"], case["E"], case["occupations"], tei_ref, "partial",
                       [{"ao_integrals": Or, "is_imaginary": False}, {"ao_integrals": Oi, "is_imaginary": True}],
                       freqs)
    for f in range(len(freqs)):
        assert np.abs(driver.results[f] - ref["results"][f]).max() <= 1e-9, (freqs, f)
        for k in range(2):
            assert np.abs(ops[k].rspvecs_alph[f] - ref["rspvecs_alph"][k][f]).max() <= 1e-9
    if any(w != 0.0 for w in freqs):
        assert solver.solve_stats.get("potri_shared", 0) == 1 and solver.solve_stats["potri"] == 0
        assert solver.solve_stats.get("potrf_shared", 0) == 1
eng.DISTRIBUTED_POTRF_MIN_N, eng.KINV_TREE_MAX_N = 2048, 12288     # defaults again
# nov = 512: local factor of K, shared K^-1 from the local inv(L) (recursive doubling) + this rank's
# column blocks of Z^T Z, all-gathered
nb_, ob_ = 72, 8
big = synthetic.rhf_case(nb_, ob_)
blo, bhi = pd.shard_range(nb_, rank, world)
ab = AO2MO(big["C"], big["occupations"], I=np.ascontiguousarray(big["I"][:, blo:bhi]),
           nu_range=(blo, bhi), device_results=True, reduce="scatter")
ab.perform_rhf_partial()
sb = solvers.ExactInv(big["C"], big["E"], big["occupations"])
sb.distribute_frequencies = True
sb.tei_mo = ab.tei_mo
sb.tei_mo_rows = ab.row_ranges[0]
db = cphf.CPHF(sb)
Ob = synthetic.operator_integrals(nb_, 3, False)
Oib = synthetic.operator_integrals(nb_, 3, True)
for O_, im_ in ((Ob, False), (Oib, True)):
    db.add_operator(operators.Operator(label="x", is_imaginary=im_, ao_integrals=O_))
db.set_frequencies([0.0, 0.2, 0.4])
db.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
assert sb.solve_stats.get("potri_shared", 0) == 1 and sb.solve_stats.get("potrf_shared", 0) == 0
tb = orc.ao2mo_rhf_partial(big["I"], big["C"], ob_)
rb = orc.run_cphf(big["C"], big["E"], big["occupations"], tb, "partial",
                  [{"ao_integrals": Ob, "is_imaginary": False}, {"ao_integrals": Oib, "is_imaginary": True}],
                  [0.0, 0.2, 0.4])
for f in range(3):
    assert np.abs(db.results[f] - rb["results"][f]).max() <= 1e-9, f
# one frequency above the first pole (owned by rank 1): ExactInv falls back to LU there and the
# gather still completes; ExactInvCholesky raises LinAlgError on EVERY rank (no rank left waiting)
freqs = [0.1, 1.9, 0.2]
ref = orc.run_cphf(case["C"], case["E"], case["occupations"], tei_ref, "partial",
                   [{"ao_integrals": Or, "is_imaginary": False}], freqs)
for cls in (solvers.ExactInv, solvers.ExactInvCholesky):
    solver = cls(case["C"], case["E"], case["occupations"])
    solver.distribute_frequencies = True
    solver.tei_mo = a.tei_mo
    driver = cphf.CPHF(solver)
    driver.add_operator(operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or))
    driver.set_frequencies(freqs)
    try:
        driver.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
        raised = False
    except np.linalg.LinAlgError:
        raised = True
    assert raised == (cls is solvers.ExactInvCholesky), (cls, raised, rank)
    if not raised:
        for f in range(3):
            assert np.abs(driver.results[f] - ref["results"][f]).max() <= 1e-9, f
# UHF, dense nu-sharded: perform_uhf_partial(reduce=True) -> the pyscf 6-tuple on every rank
nu_, na_, nb_ = 22, 4, 3
ucase = synthetic.uhf_case(nu_, na_, nb_)
ulo, uhi = pd.shard_range(nu_, rank, world)
ua = AO2MO(ucase["C"], ucase["occupations"], I=np.ascontiguousarray(ucase["I"][:, ulo:uhi]),
           nu_range=(ulo, uhi), device_results=True, reduce=True)
ua.perform_uhf_partial()
utei_ref = orc.ao2mo_uhf_partial(ucase["I"], ucase["C"], ucase["occupations"])
assert len(ua.tei_mo) == 6
for got, want in zip(ua.tei_mo, utei_ref):
    assert tuple(got.shape) == want.shape
    assert np.abs(got.cpu().numpy() - want).max() <= 1e-12 * np.abs(want).max()
usolver = solvers.ExactInv(ucase["C"], ucase["E"], ucase["occupations"])
usolver.distribute_frequencies = True
usolver.tei_mo = ua.tei_mo
udriver = cphf.CPHF(usolver)
Ou = synthetic.operator_integrals(nu_, 3, False)
udriver.add_operator(operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Ou))
udriver.set_frequencies([0.0, 0.12])
udriver.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
uref = orc.run_cphf(ucase["C"], ucase["E"], ucase["occupations"], utei_ref, "partial",
                    [{"ao_integrals": Ou, "is_imaginary": False}], [0.0, 0.12])
for f in range(2):
    assert np.abs(udriver.results[f] - uref["results"][f]).max() <= 1e-9
# ---- row-block-sharded Hessian: AO2MO(reduce="scatter") keeps only this rank's occupied rows of
# every block; the solver assembles its rows of M / K and all-gathers them (RHF and UHF); the
# matrix-free PCG solver applies sigma row block by row block with one all-gather per iteration
a2 = AO2MO(case["C"], case["occupations"], I=np.ascontiguousarray(case["I"][:, lo:hi]),
           nu_range=(lo, hi), device_results=True, reduce="scatter")
a2.perform_rhf_partial()
(ilo, ihi), _ = a2.row_ranges
assert (ilo, ihi) == pd.shard_range(o, rank, world)
assert np.abs(a2.tei_mo[0].cpu().numpy() - tei_ref[0][ilo:ihi]).max() <= 1e-12 * np.abs(tei_ref[0]).max()
assert np.abs(a2.tei_mo[1].cpu().numpy() - tei_ref[1][ilo:ihi]).max() <= 1e-12 * np.abs(tei_ref[1]).max()
freqs = [0.0, 0.1, 0.2]
spec = [{"ao_integrals": Or, "is_imaginary": False}, {"ao_integrals": Oi, "is_imaginary": True}]
ref = orc.run_cphf(case["C"], case["E"], case["occupations"], tei_ref, "partial", spec, freqs)
neg = orc.run_cphf(case["C"], case["E"], case["occupations"], tei_ref, "partial", spec, [-w for w in freqs])
solver = solvers.ExactInv(case["C"], case["E"], case["occupations"])
solver.distribute_frequencies = True
solver.tei_mo = a2.tei_mo
solver.tei_mo_rows = a2.row_ranges[0]
driver = cphf.CPHF(solver)
ops = [operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or),
       operators.Operator(label="angmom", is_imaginary=True, ao_integrals=Oi)]
for op in ops:
    driver.add_operator(op)
driver.set_frequencies(freqs)
driver.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
A_ref, B_ref = orc.rhf_ab(case["E"], tei_ref, "partial", o, "rpa", "singlet")
assert np.abs(solver._system.M.cpu().numpy() - (A_ref - B_ref)).max() <= 1e-12 * np.abs(A_ref).max()
assert np.abs(solver._system.K.cpu().numpy() - (A_ref + B_ref)).max() <= 1e-12 * np.abs(A_ref).max()
for f in range(3):
    assert np.abs(driver.results[f] - ref["results"][f]).max() <= 1e-9
it = solvers.IterativeLinEqSolver(case["C"], case["E"], case["occupations"], maxiter=80, conv=1e-26)
it.tei_mo = a2.tei_mo
it.tei_mo_rows = a2.row_ranges[0]
idrv = cphf.CPHF(it)
iops = [operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or),
        operators.Operator(label="angmom", is_imaginary=True, ao_integrals=Oi)]
for op in iops:
    idrv.add_operator(op)
idrv.set_frequencies(freqs)
idrv.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
assert it.iterations[-1] > 0, it.residuals
for f in range(3):
    assert np.abs(iops[0].rspvecs_alph[f] - neg["rspvecs_alph"][0][f]).max() <= 1e-10
    assert np.abs(iops[1].rspvecs_alph[f] + neg["rspvecs_alph"][1][f]).max() <= 1e-10
    assert np.abs(idrv.results[f][:3, :3] - ref["results"][f][:3, :3]).max() <= 1e-8
# UHF, row sharded (ragged: 4 alpha / 3 beta occupied over 2 ranks)
ua2 = AO2MO(ucase["C"], ucase["occupations"], I=np.ascontiguousarray(ucase["I"][:, ulo:uhi]),
            nu_range=(ulo, uhi), device_results=True, reduce="scatter")
ua2.perform_uhf_partial()
(al, ah), (bl_, bh) = ua2.row_ranges
for k, (r0, r1) in enumerate(((al, ah), (al, ah), (bl_, bh), (bl_, bh), (al, ah), (bl_, bh))):
    want = utei_ref[k][r0:r1]
    assert tuple(ua2.tei_mo[k].shape) == want.shape, k
    assert np.abs(ua2.tei_mo[k].cpu().numpy() - want).max() <= 1e-12 * np.abs(utei_ref[k]).max(), k
us2 = solvers.ExactInv(ucase["C"], ucase["E"], ucase["occupations"])
us2.distribute_frequencies = True
us2.tei_mo = ua2.tei_mo
us2.tei_mo_rows = ua2.row_ranges
ud2 = cphf.CPHF(us2)
ud2.add_operator(operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Ou))
ud2.set_frequencies([0.0, 0.12])
ud2.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
for f in range(2):
    assert np.abs(ud2.results[f] - uref["results"][f]).max() <= 1e-9
pd.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_ranks_sharded_ao2mo_and_distributed_frequencies(tmp_path):
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
    outs = [p.communicate(timeout=600)[0] for p in procs]
    report = "\n".join(f"--- rank {r} (rc={p.returncode}) ---\n{o[-3000:]}" for r, (p, o) in enumerate(zip(procs, outs)))
    for r, p in enumerate(procs):
        assert p.returncode == 0, report
        assert f"ok {r}" in outs[r]
