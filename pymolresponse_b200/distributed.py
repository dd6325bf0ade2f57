"""One process per GPU: how the hot path is sharded over the ranks of one box.

* AO ERI tensor: sharded over the second index of the first AO pair (``nu``); every rank holds
  ``I[:, nu_lo:nu_hi, :, :]`` with mu local, so the 2 n^4 o first quarter needs no reduction.  The
  finished (ia|jb) / (ij|ab) blocks are partial sums over the local nu and are completed with one
  ``all_reduce`` each (GBs, against tens of GBs for the half-transformed tensor).
* Orbital-Hessian assembly: row blocks by occupied index i (both index swaps stay rank-local).
* Solves: independent frequencies are dealt round-robin to ranks; the small result tensors and
  response vectors are all-gathered.

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the only transport; nothing here
does arithmetic.
"""

from __future__ import annotations

import os


def dist():
    import torch.distributed as d

    return d


def is_initialized() -> bool:
    d = dist()
    return d.is_available() and d.is_initialized()


def rank_world() -> tuple[int, int]:
    if is_initialized():
        return dist().get_rank(), dist().get_world_size()
    return 0, 1


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise from torchrun's environment; returns (rank, world, local_rank)."""
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist().init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of range(n) for this rank (first ranks get the extras)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def round_robin(n_items: int, rank: int, world: int) -> list[int]:
    """Indices of the items (frequencies) this rank owns."""
    return list(range(rank, n_items, world))


def owner_of(item: int, world: int) -> int:
    return item % world


def _host_staged() -> bool:
    """gloo moves CUDA tensors only for a few collectives: stage through the host with it (tests that
    run two ranks on one GPU); NCCL works on device memory directly."""
    return dist().get_backend() == "gloo"


def allreduce_blocks(tensors) -> None:
    """In-place sum over ranks of the partial MO-integral blocks."""
    if not is_initialized():
        return
    d = dist()
    for t in tensors:
        if t.is_cuda and _host_staged():
            h = t.cpu()
            d.all_reduce(h, op=d.ReduceOp.SUM)
            t.copy_(h)
        else:
            d.all_reduce(t, op=d.ReduceOp.SUM)


def gather_round_robin(local_items: list, n_items: int):
    """Inverse of :func:`round_robin`: every rank contributes the tensors of the items it owns
    (same shape for all items) and receives the full list in item order."""
    rank, world = rank_world()
    if world == 1:
        return list(local_items)
    import torch

    d = dist()
    per_rank_max = (n_items + world - 1) // world
    proto = None
    for t in local_items:
        proto = t
        break
    if proto is not None:
        dev = proto.device
    elif torch.cuda.is_available():
        # a rank that owns no item still returns tensors where the others live (the device)
        dev = torch.device("cuda", torch.cuda.current_device())
    else:
        dev = torch.device("cpu")
    shape_t = torch.zeros(8, dtype=torch.int64, device=dev)
    if proto is not None:
        shape_t[0] = proto.dim()
        for k, s in enumerate(proto.shape):
            shape_t[1 + k] = s
    if dev.type == "cuda" and _host_staged():
        h = shape_t.cpu()
        d.all_reduce(h, op=d.ReduceOp.MAX)
        shape_t.copy_(h)
    else:
        d.all_reduce(shape_t, op=d.ReduceOp.MAX)
    ndim = int(shape_t[0].item())
    shape = [int(shape_t[1 + k].item()) for k in range(ndim)]
    device = shape_t.device
    coll_dev = torch.device("cpu") if (device.type == "cuda" and _host_staged()) else device
    buf = torch.zeros([per_rank_max] + shape, dtype=torch.float64, device=coll_dev)
    for k, t in enumerate(local_items):
        buf[k].copy_(t)
    gathered = [torch.zeros_like(buf) for _ in range(world)]
    d.all_gather(gathered, buf)
    out = []
    for item in range(n_items):
        out.append(gathered[owner_of(item, world)][item // world].to(device).clone())
    return out


def broadcast(t, src: int) -> None:
    """In-place broadcast of a contiguous tensor from rank ``src``."""
    if not is_initialized():
        return
    d = dist()
    if t.is_cuda and _host_staged():
        h = t.cpu()
        d.broadcast(h, src=src)
        if d.get_rank() != src:
            t.copy_(h)
    else:
        d.broadcast(t, src=src)


def reduce_scatter_rows(T, n_rows_total: int):
    """Sum T (n_rows_total, ...) over ranks and return this rank's row block
    ``shard_range(n_rows_total, rank, world)`` of the sum."""
    rank, world = rank_world()
    if world == 1:
        return T
    import torch

    d = dist()
    lo, hi = shard_range(n_rows_total, rank, world)
    if T.is_cuda and not _host_staged() and n_rows_total % world == 0:
        out = torch.empty((hi - lo,) + tuple(T.shape[1:]), dtype=T.dtype, device=T.device)
        d.reduce_scatter_tensor(out, T.contiguous(), op=d.ReduceOp.SUM)
        return out
    allreduce_blocks([T])
    return T[lo:hi].contiguous()


def allgather_rows(X, n_rows_total: int):
    """Inverse of the row sharding: every rank contributes its row block, all get the whole."""
    rank, world = rank_world()
    if world == 1:
        return X
    import torch

    d = dist()
    rmax = (n_rows_total + world - 1) // world
    coll_dev = torch.device("cpu") if (X.is_cuda and _host_staged()) else X.device
    buf = torch.zeros((rmax,) + tuple(X.shape[1:]), dtype=X.dtype, device=coll_dev)
    buf[: X.shape[0]].copy_(X)
    if coll_dev == X.device and n_rows_total % world == 0:
        out = torch.empty((n_rows_total,) + tuple(X.shape[1:]), dtype=X.dtype, device=X.device)
        d.all_gather_into_tensor(out, buf)
        return out
    parts = [torch.empty_like(buf) for _ in range(world)]
    d.all_gather(parts, buf)
    out = torch.empty((n_rows_total,) + tuple(X.shape[1:]), dtype=X.dtype, device=X.device)
    for r in range(world):
        lo, hi = shard_range(n_rows_total, r, world)
        out[lo:hi].copy_(parts[r][: hi - lo])
    return out


def allgather_columns(X, n_cols_total: int):
    """Every rank holds the column block ``shard_range(n_cols_total, rank, world)`` of a matrix as a
    contiguous (rows, width) tensor; returns the full (rows, n_cols_total) matrix on every rank."""
    rank, world = rank_world()
    if world == 1:
        return X
    import torch

    d = dist()
    rows = int(X.shape[0])
    wmax = (n_cols_total + world - 1) // world
    coll_dev = torch.device("cpu") if (X.is_cuda and _host_staged()) else X.device
    if int(X.shape[1]) == wmax and X.is_contiguous() and coll_dev == X.device:
        buf = X
    else:
        buf = torch.zeros(rows, wmax, dtype=X.dtype, device=coll_dev)
        buf[:, : X.shape[1]].copy_(X)
    out = torch.empty(rows, n_cols_total, dtype=X.dtype, device=X.device)
    if coll_dev == X.device:
        # one flat all-gather (world, rows, wmax), then the column blocks go to their places
        flat = torch.empty(world * rows, wmax, dtype=X.dtype, device=coll_dev)
        d.all_gather_into_tensor(flat, buf)           # rank blocks concatenated along dim 0
        parts = [flat[r * rows:(r + 1) * rows] for r in range(world)]
    else:
        parts = [torch.empty_like(buf) for _ in range(world)]
        d.all_gather(parts, buf)
    for r in range(world):
        lo, hi = shard_range(n_cols_total, r, world)
        out[:, lo:hi].copy_(parts[r][:, : hi - lo])
    return out


def any_rank(flag: bool) -> bool:
    """True on every rank if ``flag`` is true on at least one (collective agreement on a failure
    before the next data-path collective)."""
    if not is_initialized():
        return bool(flag)
    import torch

    d = dist()
    if d.get_backend() == "gloo" or not torch.cuda.is_available():
        t = torch.tensor([1 if flag else 0], dtype=torch.int32)
    else:
        t = torch.tensor([1 if flag else 0], dtype=torch.int32,
                         device=torch.device("cuda", torch.cuda.current_device()))
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return bool(int(t.item()))


def barrier() -> None:
    if is_initialized():
        dist().barrier()
