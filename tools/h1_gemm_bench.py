"""First-quarter-shaped GEMM of the AO->MO transform: H1[i, x] = sum_mu Co[mu, i] I[mu, x] with x a
long contiguous run (a nu-block of the slab).  Times libpmr_b200 against cuBLAS (torch.matmul).
Usage: python tools/h1_gemm_bench.py [n o nb]   (env PMR_GEMM_NO_BULK / PMR_GEMM_TILE select kernels)"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from pymolresponse_b200 import _native as nv  # noqa: E402

n, o, nb = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (512, 64, 8)
X = nb * n * n
I = torch.randn(n, X, dtype=torch.float64, device="cuda")
Co = torch.randn(n, o, dtype=torch.float64, device="cuda")
H = torch.empty(o, X, dtype=torch.float64, device="cuda")


def timed(fn, reps=5):
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


ms = timed(lambda: nv.dgemm(o, X, n, 1.0, Co, 0, 1, o, I, 0, X, 1, 0.0, H, 0, X, 1))
ref = Co.t() @ I
err = float((H - ref).abs().max())
ms_ref = timed(lambda: torch.matmul(Co.t(), I, out=ref))
fl = 2.0 * o * X * n
print(json.dumps({"shape": [o, X, n], "env": {k: v for k, v in os.environ.items() if k.startswith("PMR_")},
                  "ms": ms, "tflops": fl / ms / 1e9, "GBps_read": I.numel() * 8 / ms / 1e6,
                  "cublas_ms": ms_ref, "cublas_tflops": fl / ms_ref / 1e9, "max_err": err}))
