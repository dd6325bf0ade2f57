"""Trailing-update-shaped GEMM, C (rows x rows) -= P P^T with P = (rows x 512), batched over 64
matrices of leading dimension 7168: full / lower-triangular, beta 0 / 1.  Development aid."""
import json
import sys

import torch

sys.path.insert(0, ".")
from pymolresponse_b200 import _native as nv  # noqa: E402

N, K, nb = 7168, 512, 64
rows_list = [int(a) for a in sys.argv[1:]] or [6656, 3584, 1024]
S = torch.randn(nb, N, N, dtype=torch.float64, device="cuda")


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for rows in rows_list:
    r0 = N - rows
    j0 = r0 - K
    for lower in (0, 1):
        for beta in (0.0, 1.0):
            def run():
                nv.dgemm(rows, rows, K, -1.0, S, r0 * N + j0, N, 1, S, r0 * N + j0, 1, N, beta,
                         S, r0 * N + r0, N, 1, batch1=nb, a_b=(N * N, 0), b_b=(N * N, 0),
                         c_b=(N * N, 0), flags=nv.GEMM_LOWER if lower else 0)
            ms = timed(run)
            mn = rows * (rows + 1) / 2 if lower else rows * rows
            print(json.dumps({"rows": rows, "lower": lower, "beta": beta, "ms": round(ms, 3),
                              "tflops": round(2.0 * mn * K * nb / ms / 1e9, 2)}), flush=True)
