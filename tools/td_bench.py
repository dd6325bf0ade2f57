"""Time the excitation-energy path on the GPU and the reference's algorithm
(scipy.linalg.eig on the 2N x 2N matrix, via the oracle) on the host, same synthetic inputs.
Usage: python tools/td_bench.py [n nocc] [--no-cpu]  ->  one JSON line"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pymolresponse_b200 import _native as nv, operators, solvers, synthetic, td  # noqa: E402
from pymolresponse_b200.ao2mo import AO2MO  # noqa: E402
from pymolresponse_b200.core import AO2MOTransformationType, Hamiltonian, Spin  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n, o = (int(args[0]), int(args[1])) if len(args) >= 2 else (64, 12)
case = synthetic.rhf_case(n, o, build_eri=False)
a = AO2MO(case["C"], case["occupations"], df_factor=case["B"], device_results=True)
a.perform_rhf_partial()
tei = a.tei_mo
Or, Oi = synthetic.operator_integrals(n, 3, False), synthetic.operator_integrals(n, 3, True)
nov = o * (n - o)


def run():
    s = solvers.ExactDiagonalizationSolver(case["C"], case["E"], case["occupations"])
    s.tei_mo = tei
    s.tei_mo_type = AO2MOTransformationType.partial
    d = td.TDHF(s)
    d.add_operator(operators.Operator(label="angmom", is_imaginary=True, ao_integrals=Oi))
    d.add_operator(operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, s


run()
nv.reset_launch_count()
t_gpu, s = run()
out = {"workload": f"RPA singlet excitation energies, all {nov} roots, synthetic RHF nbasis={n} nocc={o}",
       "nov": nov, "gpu_seconds": t_gpu, "jacobi_sweeps": s.jacobi_sweeps,
       "gpu_launches": nv.launch_count(), "lowest_roots": [float(x) for x in s.eigvals[:4]]}
if "--no-cpu" not in sys.argv:
    from oracle import pmr_oracle as orc

    tei_h = tuple(nv.to_host(t) for t in tei)
    t0 = time.perf_counter()
    ev, _ = orc.diagonalize_rhf(case["E"], tei_h, "partial", o, "rpa", "singlet")
    out["cpu_seconds_reference_algorithm"] = time.perf_counter() - t0
    out["max_abs_diff_eigvals"] = float(np.abs(ev.real - s.eigvals).max())
print(json.dumps(out))
