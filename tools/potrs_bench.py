"""Time the triangular solves of the sweep in isolation: batched potrs (64 x 7168^2, 3 columns) and
the single-matrix potrs with 192 and 3 columns.  Usage: python tools/potrs_bench.py [N batch]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from pymolresponse_b200 import _engine as eng, _native as nv  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 7168
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
X = torch.randn(N, N, dtype=torch.float64, device="cuda")
A = X @ X.t() / N + 1.5 * torch.eye(N, dtype=torch.float64, device="cuda")
del X
S = A.unsqueeze(0).repeat(nb, 1, 1).contiguous()
dinv, info = eng.potrf(S, batch=nb)
assert int(info.abs().sum().item()) == 0


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


out = {}
for nrhs in (3, 6, 9):
    B = torch.randn(nb, N, nrhs, dtype=torch.float64, device="cuda")
    out[f"batched_{nb}x{N}_nrhs{nrhs}_ms"] = timed(lambda: eng.potrs(S, dinv, B.clone(), batch=nb))
    out[f"batched_nrhs{nrhs}_GBps_factor_reads"] = 2 * nb * N * N / 2 * 8 / out[f"batched_{nb}x{N}_nrhs{nrhs}_ms"] / 1e6
d1 = dinv[0:1].contiguous()
for nrhs in (3, 192):
    B = torch.randn(N, nrhs, dtype=torch.float64, device="cuda")
    out[f"single_{N}_nrhs{nrhs}_ms"] = timed(lambda: eng.potrs(S[0], d1, B.clone()))
print(json.dumps(out))
