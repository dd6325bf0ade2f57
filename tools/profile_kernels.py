"""Small driver for `ncu --set full` captures of the three kernel families (development aid):
trailing-update GEMM (cp.async, lower), TMA-bulk GEMM (first AO->MO quarter shape), fast assembly."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pymolresponse_b200 import _blocks as bl  # noqa: E402
from pymolresponse_b200 import _native as nv  # noqa: E402

which = sys.argv[1]
if which == "trailing":      # one rank-512 lower-triangular update of 16 x 7168^2 matrices
    nb, N, K = 16, 6656, 512
    P = torch.randn(nb, N, K, dtype=torch.float64, device="cuda")
    Cm = torch.randn(nb, N, N, dtype=torch.float64, device="cuda")
    for _ in range(2):
        nv.dgemm(N, N, K, -1.0, P, 0, K, 1, P, 0, 1, K, 1.0, Cm, 0, N, 1, batch1=nb,
                 a_b=(N * K, 0), b_b=(N * K, 0), c_b=(N * N, 0), flags=nv.GEMM_LOWER)
elif which == "bulk":        # first AO->MO quarter: (n^2 cnt) x o x n with the AO tensor as streamed A
    n, o, cnt = 256, 32, 64
    I = torch.randn(n, cnt * n * n, dtype=torch.float64, device="cuda")
    Co = torch.randn(n, o, dtype=torch.float64, device="cuda")
    H = torch.empty(o, cnt * n * n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        nv.dgemm(o, cnt * n * n, n, 1.0, Co, 0, 1, o, I, 0, cnt * n * n, 1, 0.0, H, 0, cnt * n * n, 1)
elif which == "assemble":
    o, v = 64, 448
    nov = o * v
    ovov = torch.randn(o, v, o, v, dtype=torch.float64, device="cuda")
    oovv = torch.randn(o, o, v, v, dtype=torch.float64, device="cuda")
    eps = (torch.linspace(-2, -0.5, o, dtype=torch.float64, device="cuda"),
           torch.linspace(0.5, 3, v, dtype=torch.float64, device="cuda"))
    M, K = nv.empty(nov, nov), nv.empty(nov, nov)
    for _ in range(2):
        bl.run(o, v, o, v, bl.partial_terms("singlet", ovov, oovv), eps, "PQ", out_mode=1, out0=M,
               out1=K, ld=nov)
torch.cuda.synchronize()
print("ok")
