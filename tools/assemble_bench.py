"""Hessian-assembly kernel on its own: achieved GB/s (algorithmic bytes 4*nov^2*8) vs HBM peak."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pymolresponse_b200 import _blocks as bl  # noqa: E402
from pymolresponse_b200 import _native as nv  # noqa: E402

o = int(sys.argv[1]) if len(sys.argv) > 1 else 32
v = int(sys.argv[2]) if len(sys.argv) > 2 else 224
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
nov = o * v
ovov = torch.randn(o, v, o, v, dtype=torch.float64, device="cuda")
oovv = torch.randn(o, o, v, v, dtype=torch.float64, device="cuda")
eps = (torch.linspace(-2, -0.5, o, dtype=torch.float64, device="cuda"),
       torch.linspace(0.5, 3, v, dtype=torch.float64, device="cuda"))
M, K = nv.empty(nov, nov), nv.empty(nov, nov)
terms = bl.partial_terms("singlet", ovov, oovv)
for mode in (1, 0):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bl.run(o, v, o, v, terms, eps, "PQ", out_mode=mode, out0=M, out1=K, ld=nov)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps({"kernel": "assemble", "out_mode": mode, "o": o, "v": v, "nov": nov, "ms": best,
                      "GBps": 4.0 * nov * nov * 8 / best / 1e6}))
