"""Phase timing of the first cluster-resident inner panel of pmr_getrf (clock64 stamps of thread 0 of
CTA 0 through pmr_debug_potf2_timing).  Development aid."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pymolresponse_b200 import _engine as eng  # noqa: E402
from pymolresponse_b200 import _native as nv  # noqa: E402

torch.manual_seed(0)
for N in [int(a) for a in sys.argv[1:]] or [14336, 2048]:
    A = torch.randn(N, N, dtype=torch.float64, device="cuda")
    buf = torch.zeros(64, dtype=torch.int64, device="cuda")
    for rep in range(2):
        G = A.clone()
        nv.check(nv.lib().pmr_debug_potf2_timing(nv.ptr(buf)))
        eng.getrf(G)
        torch.cuda.synchronize()
        nv.check(nv.lib().pmr_debug_potf2_timing(None))
    t = buf.cpu().tolist()
    print(f"N = {N}: first inner panel {t[44] - t[0]} cycles")
    print(f"  load {t[1] - t[0]}")
    for jj in range(8):
        b = 2 + 5 * jj
        prev = t[1] if jj == 0 else t[b - 1]
        print(f"  column {jj}: candidate+reduce {t[b] - prev}  publish {t[b + 1] - t[b]}  "
              f"cluster barrier {t[b + 2] - t[b + 1]}  swap+sync {t[b + 3] - t[b + 2]}  update {t[b + 4] - t[b + 3]}")
    print(f"  flush {t[42] - t[41]}  store {t[43] - t[42]}  U12 {t[44] - t[43]}")
