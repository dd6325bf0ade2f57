"""Per-stage device timings (CUDA events) of the hot path + GEMM microbenchmarks vs cuBLAS DGEMM.
Development aid; writes JSON lines to stdout."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pymolresponse_b200 import _blocks as bl  # noqa: E402
from pymolresponse_b200 import _engine as eng  # noqa: E402
from pymolresponse_b200 import _native as nv  # noqa: E402
from pymolresponse_b200 import synthetic  # noqa: E402


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def gemm_micro():
    out = []
    shapes = [(8192, 8192, 8192, "kc", "kn"), (8192, 8192, 128, "kc", "kc"), (8192, 8192, 512, "kc", "kc"),
              (65536, 32, 256, "mc", "kn"), (16384, 128, 128, "kc", "kc"), (4096, 4096, 4096, "mc", "kn")]
    for M, N, K, la, lb in shapes:
        A = torch.randn(M, K, dtype=torch.float64, device="cuda") if la == "kc" else torch.randn(K, M, dtype=torch.float64, device="cuda")
        B = torch.randn(N, K, dtype=torch.float64, device="cuda") if lb == "kc" else torch.randn(K, N, dtype=torch.float64, device="cuda")
        Cm = torch.zeros(M, N, dtype=torch.float64, device="cuda")
        a_sm, a_sk = (K, 1) if la == "kc" else (1, M)
        b_sk, b_sn = (1, K) if lb == "kc" else (N, 1)
        ms = timed(lambda: nv.dgemm(M, N, K, 1.0, A, 0, a_sm, a_sk, B, 0, b_sk, b_sn, 0.0, Cm, 0, N, 1))
        Aop = A if la == "kc" else A.t()
        Bop = B.t() if lb == "kc" else B
        ms_ref = timed(lambda: torch.matmul(Aop, Bop))
        err = float((Cm - torch.matmul(Aop, Bop)).abs().max())
        fl = 2.0 * M * N * K
        out.append({"gemm": [M, N, K, la, lb], "ms": ms, "tflops": fl / ms / 1e9, "cublas_ms": ms_ref,
                    "cublas_tflops": fl / ms_ref / 1e9, "max_err": err})
        print(json.dumps(out[-1]), flush=True)
        del A, B, Cm
    return out


def stages(n, o, nf):
    v = n - o
    nov = o * v
    case = synthetic.rhf_case(n, o, build_eri=False)
    B = torch.from_numpy(case["B"]).cuda()
    Bf = B.reshape(n, n * n)
    I = torch.empty(n, n, n, n, dtype=torch.float64, device="cuda")
    for mu in range(n):
        I[mu] = (B[:, mu, :].t() @ Bf).reshape(n, n, n)
    del B, Bf
    from pymolresponse_b200.ao2mo import AO2MO
    a = AO2MO(case["C"], case["occupations"], I=I, device_results=True)
    ms_ao = timed(a.perform_rhf_partial, reps=2)
    ovov, oovv = a.tei_mo
    print(json.dumps({"stage": "ao2mo", "n": n, "o": o, "ms": ms_ao,
                      "tflops": synthetic.flops_ao2mo_rhf_partial(n, o) / ms_ao / 1e9}), flush=True)
    del I
    torch.cuda.empty_cache()
    eps = bl.split_energies(case["E"][0], o)
    M, K = nv.empty(nov, nov), nv.empty(nov, nov)
    terms = bl.partial_terms("singlet", ovov, oovv)
    ms_as = timed(lambda: bl.run(o, v, o, v, terms, eps, "PQ", out_mode=1, out0=M, out1=K, ld=nov))
    print(json.dumps({"stage": "assemble", "nov": nov, "ms": ms_as,
                      "GBps": synthetic.bytes_assemble(nov) / ms_as / 1e6}), flush=True)
    u = nov ** 3 / 3.0
    Kf = K.clone()
    def f_potrf():
        Kf.copy_(K)
        return eng.potrf(Kf)
    ms_copy = timed(lambda: Kf.copy_(K))
    ms_p = timed(f_potrf) - ms_copy
    dinv, info = f_potrf()
    print(json.dumps({"stage": "potrf(K) single", "ms": ms_p, "tflops": u / ms_p / 1e9,
                      "info": int(info.item())}), flush=True)
    ms_i = timed(lambda: eng.potri_lower(Kf, dinv))
    Kinv = eng.potri_lower(Kf, dinv)
    print(json.dumps({"stage": "potri_lower", "ms": ms_i, "tflops": 2 * u / ms_i / 1e9}), flush=True)
    ws = list(np.linspace(0.0, 0.63, nf))
    ms_c = timed(lambda: eng.shifted_combine(M, Kinv, [0.0] * nf, [-(w * w) for w in ws]))
    S = eng.shifted_combine(M, Kinv, [0.0] * nf, [-(w * w) for w in ws])
    print(json.dumps({"stage": "shifted_combine", "nf": nf, "ms": ms_c,
                      "GBps": (2 + nf) * nov * nov * 4 / ms_c / 1e6}), flush=True)
    S0 = S.clone()
    def f_bp():
        S.copy_(S0)
        return eng.potrf(S, batch=nf)
    ms_copy = timed(lambda: S.copy_(S0))
    ms_b = timed(f_bp, reps=2) - ms_copy
    dinvS, infoS = f_bp()
    print(json.dumps({"stage": "potrf batched", "nf": nf, "ms": ms_b, "tflops": nf * u / ms_b / 1e9,
                      "bad": int((infoS != 0).sum().item())}), flush=True)
    rhs = torch.randn(nf, nov, 3, dtype=torch.float64, device="cuda")
    ms_s = timed(lambda: eng.potrs(S, dinvS, rhs.clone(), batch=nf))
    print(json.dumps({"stage": "potrs batched nrhs=3", "ms": ms_s,
                      "GBps": nf * nov * nov * 8 / ms_s / 1e6}), flush=True)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    o = int(sys.argv[2]) if len(sys.argv) > 2 else n // 8
    nf = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    gemm_micro()
    stages(n, o, nf)
