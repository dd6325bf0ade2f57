"""One single-matrix and one batched Cholesky, for ncu launch lists (development aid)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pymolresponse_b200 import _engine as eng  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4032
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 8
X = torch.randn(N, N, dtype=torch.float64, device="cuda")
A = X @ X.t() / N + 1.5 * torch.eye(N, dtype=torch.float64, device="cuda")
for _ in range(2):
    K = A.clone()
    dinv, info = eng.potrf(K)
torch.cuda.synchronize()
S = A.unsqueeze(0).repeat(nf, 1, 1).contiguous()
dinv, info = eng.potrf(S, batch=nf)
torch.cuda.synchronize()
print("ok", int(info.sum().item()))
