"""Why are the all-gathers of the stage-clock pass slow?  Times torch.distributed all-gathers of the
config-5 K^-1 size under a few conditions (development aid; run under torchrun)."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pymolresponse_b200 import distributed as pd  # noqa: E402

rank, world, local = pd.init_from_env()
torch.cuda.set_device(local)
N = 7168
X = torch.randn(N, N // world, dtype=torch.float64, device="cuda")


def timed(fn, label, pre_sync=False, fresh=False):
    ts = []
    for _ in range(6):
        if pre_sync:
            torch.cuda.synchronize()
        if fresh:
            torch.cuda.empty_cache()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    if rank == 0:
        print(f"{label:50s} " + " ".join(f"{t:7.2f}" for t in ts), flush=True)


def flat():
    out = torch.empty(world, N, N // world, dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(out, X)
    return out


def listed():
    parts = [torch.empty_like(X) for _ in range(world)]
    dist.all_gather(parts, X)
    return parts


timed(flat, "all_gather_into_tensor")
timed(listed, "all_gather(list)")
timed(flat, "all_gather_into_tensor, device sync before", pre_sync=True)
timed(listed, "all_gather(list), device sync before", pre_sync=True)
timed(flat, "all_gather_into_tensor, empty_cache before", pre_sync=True, fresh=True)
timed(lambda: pd.allgather_columns(X, N), "pd.allgather_columns")
Y = torch.randn(N // world, N, dtype=torch.float64, device="cuda")
timed(lambda: pd.allgather_rows(Y, N), "pd.allgather_rows")
dist.barrier()
dist.destroy_process_group()
