"""Time the AO ERI uploads (whole array, s8 row-prefix, s8 box) from pinned host memory.
Usage: python tools/upload_bench.py [n]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from pymolresponse_b200 import _native as nv  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = nv.lib()
host = torch.empty((n, n, n, n), dtype=torch.float64, pin_memory=True)
host.zero_()
arr = host.numpy()
dev = nv.empty(n, n, n, n)
st = nv.stream_ptr()


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


t = timed(lambda: dev.copy_(host))
print(f"whole array: {t:8.1f} ms  {arr.nbytes / t / 1e6:6.1f} GB/s")
import ctypes as C  # noqa: E402

for layout, ns in ((0, 1), (1, 1), (2, 1)):
    nb = int(L.pmr_eri_s8_bytes(n, layout))
    tu = timed(lambda: nv.check(L.pmr_eri_upload_s8(C.c_void_p(arr.ctypes.data), nv.ptr(dev), n, layout, st)))
    tc = timed(lambda: nv.check(L.pmr_eri_complete_s8(nv.ptr(dev), n, layout, st)))
    print(f"layout {layout} streams {ns}: upload {tu:8.1f} ms ({nb / 1e9:.2f} GB, {nb / tu / 1e6:6.1f} GB/s)  "
          f"complete {tc:6.1f} ms")
