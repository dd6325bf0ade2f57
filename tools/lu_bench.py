"""Time pmr_getrf + pmr_getrs (pivoted LU on the full super-Hessian, the above-the-pole path) at
the config-5 size 2 nov = 14336 and smaller, against the algorithmic 2/3 N^3 (development aid)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pymolresponse_b200 import _engine as eng  # noqa: E402
from pymolresponse_b200 import _native as nv  # noqa: E402

for N in (2048, 4096, 8192, 14336):
    X = torch.randn(N, N, dtype=torch.float64, device="cuda")
    A = X + N ** 0.5 * torch.eye(N, dtype=torch.float64, device="cuda")
    B = torch.randn(N, 6, dtype=torch.float64, device="cuda")
    best = 1e9
    for rep in range(2):
        G = A.clone()
        nv.reset_launch_count()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        ipiv, info = eng.getrf(G)
        e1.record()
        Xs = eng.getrs(G, ipiv, B.clone())
        e2.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    resid = float((A @ Xs - B).abs().max())
    print(json.dumps({"N": N, "getrf_ms": round(best, 2), "getrs_ms": round(e1.elapsed_time(e2), 2),
                      "tflops": round(2.0 / 3.0 * N**3 / best / 1e9, 2), "launches": nv.launch_count(),
                      "residual": resid, "info": int(info.item())}), flush=True)
