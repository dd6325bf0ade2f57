"""Per-step wall times of the bench step (config 5, one GPU) with look-ahead on / off."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from pymolresponse_b200 import _native as nv  # noqa: E402


class A:
    n = 0; nocc = 0; nfreq = 0; mag = 0


w = bench.workload_for(1, A())
inp = bench.build_inputs(w, 0, 1, torch)
lib = nv.lib()
for la in (1, 0, 1, 0):
    lib.pmr_set_lookahead(la)
    times = []
    for k in range(7):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = {}
        bench.one_step(w, inp, inp["I"], True, 1, stage_ms=st if k == 6 else None)
        e1.record()
        torch.cuda.synchronize()
        times.append(round(e0.elapsed_time(e1), 1))
    print("lookahead", la, times, {k: round(v, 1) for k, v in st.items()}, flush=True)
