"""Phase timing of the Cholesky diagonal-block kernel (clock64 stamps written by thread 0 through
pmr_debug_potf2_timing): one 128 x 128 block of a 2048^2 matrix.  Development aid."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pymolresponse_b200 import _engine as eng  # noqa: E402
from pymolresponse_b200 import _native as nv  # noqa: E402

N = 7168
X = torch.randn(N, 256, dtype=torch.float64, device="cuda")
A = X @ X.t() / 256 + 2.0 * torch.eye(N, dtype=torch.float64, device="cuda")
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
for rep in range(2):
    K = A[:128, :128].clone()  # a single 128-block: exactly one diagonal-block launch
    Kp = torch.zeros(128, N, dtype=torch.float64, device="cuda")
    Kp[:, :128] = K
    nv.check(nv.lib().pmr_debug_potf2_timing(nv.ptr(buf)))
    eng.potrf(Kp[:, :128].clone() if rep == 0 else Kp.as_strided((128, 128), (N, 1)))
    torch.cuda.synchronize()
    nv.check(nv.lib().pmr_debug_potf2_timing(None))
    t = buf.cpu().tolist()
    names = ["load"] + [f"kb{k}:{p}" for k in range(4) for p in ("diag+inv", "X=P invD", "syrk")] + \
            ["store L", "inv j=2", "inv j=1", "inv j=0", "store dinv"]
    print(f"lda = {128 if rep == 0 else N}: total {(t[18] - t[0])} cycles")
    for k, name in enumerate(names):
        print(f"  {name:12s} {t[k + 1] - t[k]:8d}")
