"""BASELINE config 4 (UHF, nbasis 384, 48 alpha / 47 beta electrons, static polarizability) at FULL
size on ONE GPU through the factorised-integral route (AO2MO df_factor): the dense AO tensor would
be 174 GB.  Prints one JSON line with stage times.  Usage: python tools/config4_df.py [n na nb]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pymolresponse_b200 import cphf, operators, solvers, synthetic  # noqa: E402
from pymolresponse_b200.ao2mo import AO2MO  # noqa: E402
from pymolresponse_b200.core import Hamiltonian, Spin  # noqa: E402

args = [int(a) for a in sys.argv[1:]]
n, na, nb = args if len(args) == 3 else (384, 48, 47)
case = synthetic.uhf_case(n, na, nb, build_eri=False)
Or = synthetic.operator_integrals(n, 3, False)


def step():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    a = AO2MO(case["C"], case["occupations"], df_factor=case["B"], device_results=True)
    a.perform_uhf_partial()
    ev[1].record()
    s = solvers.ExactInv(case["C"], case["E"], case["occupations"])
    s.tei_mo = a.tei_mo
    del a
    d = cphf.CPHF(s)
    d.add_operator(operators.Operator(label="dipole", is_imaginary=False, ao_integrals=Or))
    d.set_frequencies([0.0])
    d.run(hamiltonian=Hamiltonian.RPA, spin=Spin.singlet, program=None, program_obj=None)
    ev[2].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), d.results[0], getattr(s, "solve_stats", {})


step()
t_ao, t_solve, alpha, stats = step()
nova, novb = na * (n - na), nb * (n - nb)
flops_solve = (nova + novb) ** 3 / 3.0
print(json.dumps({
    "workload": f"config 4: synthetic UHF nbasis={n} ({na},{nb}), static polarizability, factorised AO integrals",
    "nov": [nova, novb], "ao2mo_ms": t_ao, "assemble_solve_results_ms": t_solve,
    "total_ms": t_ao + t_solve, "solve_tflops_algorithmic": flops_solve / (t_solve * 1e-3) / 1e12,
    "alpha_diag": [float(x) for x in np.diag(alpha)],
    "alpha_asymmetry": float(np.abs(alpha - alpha.T).max()), "solve_stats": stats}))
