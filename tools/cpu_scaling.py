"""Per-stage timings of the reference's algorithm (oracle port) on the host at growing basis sizes,
power-law fits and the extrapolation to the BASELINE configurations the reference itself cannot run
on a host (n = 512 needs >= 655 GB for its intermediates).  CPU only.

    python tools/cpu_scaling.py [--sizes 48,64,80,96,112,128] > profiles/r01_cpu_extrapolation.json
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import pmr_oracle as orc  # noqa: E402
from pymolresponse_b200 import synthetic  # noqa: E402

sizes = [48, 64, 80, 96, 112, 128]
for i, a in enumerate(sys.argv):
    if a == "--sizes":
        sizes = [int(x) for x in sys.argv[i + 1].split(",")]


def best(fn, reps=2):
    t = 1e30
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        t = min(t, time.perf_counter() - t0)
    return t, out


rows = []
for n in sizes:
    o = max(2, n // 8)
    case = synthetic.rhf_case(n, o)
    Or = synthetic.operator_integrals(n, 3, False)
    t_ao, tei = best(lambda: orc.ao2mo_rhf_partial(case["I"], case["C"], o), reps=1 if n > 96 else 2)
    t_h, G = best(lambda: orc.explicit_hessian_rhf(case["E"], tei, "partial", o, "rpa", "singlet", 0.1))
    t_inv, Ginv = best(lambda: np.linalg.inv(G))
    sv = orc.form_rhs(Or, case["C"], case["occupations"])["mo_integrals_ai_supervector_alph"]
    t_res, _ = best(lambda: [np.dot(Ginv, sv[c]) for c in range(3)])
    rows.append({"n": n, "o": o, "nov": o * (n - o), "ao2mo_s": t_ao, "hessian_s": t_h,
                 "inverse_per_frequency_s": t_inv, "solve_and_results_s": t_res})
    print(json.dumps(rows[-1]), file=sys.stderr, flush=True)


def fit(xs, ys):
    """log y = a + p log x  ->  (exp(a), p)"""
    p, a = np.polyfit(np.log(xs), np.log(ys), 1)
    return float(np.exp(a)), float(p)


ns = np.array([r["n"] for r in rows], dtype=float)
novs = np.array([r["nov"] for r in rows], dtype=float)
fits = {
    "ao2mo_s ~ c n^p (o = n/8)": fit(ns, [r["ao2mo_s"] for r in rows]),
    "hessian_s ~ c nov^p": fit(novs, [r["hessian_s"] for r in rows]),
    "inverse_per_frequency_s ~ c nov^p": fit(novs, [r["inverse_per_frequency_s"] for r in rows]),
}


def extrapolate(n, o, nfreq):
    nov = o * (n - o)
    c, p = fits["ao2mo_s ~ c n^p (o = n/8)"]
    ao = c * n**p
    c, p = fits["hessian_s ~ c nov^p"]
    he = c * nov**p * nfreq
    c, p = fits["inverse_per_frequency_s ~ c nov^p"]
    inv = c * nov**p * nfreq
    return {"n": n, "o": o, "nov": nov, "frequencies": nfreq, "ao2mo_s": ao, "hessian_s": he,
            "inverse_s": inv, "total_s": ao + he + inv,
            "memory_note": f"explicit (2nov)^2 matrix and its inverse: {2 * (2 * nov) ** 2 * 8 / 1e9:.0f} GB; "
                           f"AO tensor {n ** 4 * 8 / 1e9:.0f} GB"}


print(json.dumps({
    "what": "reference algorithm (oracle port: einsum quarter transforms, explicit (2nov)^2 Hessian, "
            "np.linalg.inv per frequency) timed per stage on this host, power-law fits, extrapolation",
    "host": {"cores": os.cpu_count(), "blas": "OpenBLAS (numpy default threads)"},
    "measured": rows,
    "fits (coefficient, exponent)": fits,
    "extrapolated": {"config 5 (n=256, o=32, 64 frequencies)": extrapolate(256, 32, 64),
                     "config 3 (n=512, o=64, 8 frequencies)": extrapolate(512, 64, 8)},
    "label": "EXTRAPOLATED, not measured: the reference cannot hold these sizes on a host",
}, indent=1))
