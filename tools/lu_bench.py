"""Time pmr_getrf + pmr_getrs (pivoted LU on the full super-Hessian, the above-the-pole path) at
the config-5 size 2 nov = 14336 and smaller, against the algorithmic 2/3 N^3 (development aid)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pymolresponse_b200 import _engine as eng  # noqa: E402
from pymolresponse_b200 import _native as nv  # noqa: E402

torch.manual_seed(0)
SIZES = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192, 14336]
for N in SIZES:
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
    getrs_ms = e1.elapsed_time(e2)
    # yardstick: the vendor library's getrf through torch (not used by the product path)
    lib_ms = 1e9
    for rep in range(2):
        G = A.clone()
        e0.record()
        LUlib, pivlib = torch.linalg.lu_factor(G)
        e1.record()
        torch.cuda.synchronize()
        lib_ms = min(lib_ms, e0.elapsed_time(e1))
    print(json.dumps({"N": N, "getrf_ms": round(best, 2), "getrs_ms": round(getrs_ms, 2),
                      "tflops": round(2.0 / 3.0 * N**3 / best / 1e9, 2), "launches": nv.launch_count(),
                      "residual": resid, "library_residual": float((A @ torch.linalg.lu_solve(LUlib, pivlib, B) - B).abs().max()), "library_getrf_ms": round(lib_ms, 2), "info": int(info.item())}), flush=True)
