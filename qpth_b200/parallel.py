"""Batch sharding of the QP layer over `torch.distributed` ranks (one process per GPU).

QPs are independent, so multi-GPU = contiguous split of the batch dimension; the only
collectives are the plumbing either side of the solve:

  * `scatter_batch` / `gather_batch`: rank `src` holds the global tensors, every rank gets / returns its shard
    (NCCL over NVLink on GPUs, gloo in the CPU tests);
  * `allreduce_shared_grad`: the reference averages the gradients of un-batched inputs over the batch
    (`.mean(0)`, qpth/qp.py:159-177); with shards of unequal size the global mean is the shard-size-weighted
    mean of the per-shard means.

When the inputs are already produced data-parallel (the normal OptNet case) none of this is needed: each rank
calls `QPFunction` on its own batch and DDP reduces the caller's parameter gradients.
"""
import torch
import torch.distributed as dist


def shard_bounds(nbatch, world_size, rank):
    """Contiguous split: the first `nbatch % world_size` ranks get one extra QP."""
    base, rem = divmod(nbatch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def scatter_batch(t, nbatch, src=0, group=None, device=None):
    """Rank `src` passes the global tensor (leading dim nbatch), others pass a tensor giving dtype/trailing shape
    or None together with `like`. Returns this rank's shard. Un-batched tensors should simply be broadcast."""
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    meta = [None]
    if rk == src:
        meta = [(tuple(t.shape[1:]), t.dtype)]
    dist.broadcast_object_list(meta, src=src, group=group)
    trail, dtype = meta[0]
    lo, hi = shard_bounds(nbatch, ws, rk)
    dev = device if device is not None else (t.device if t is not None else torch.device("cpu"))
    out = torch.empty((hi - lo,) + trail, dtype=dtype, device=dev)
    chunks = None
    if rk == src:
        chunks = [t[slice(*shard_bounds(nbatch, ws, r))].contiguous().to(dev) for r in range(ws)]
    # ragged shards: scatter needs equal sizes, so send point to point
    if rk == src:
        reqs = []
        for r in range(ws):
            if r == src:
                out.copy_(chunks[r])
            elif chunks[r].numel() > 0:
                reqs.append(dist.isend(chunks[r], dst=r, group=group))
        for q in reqs:
            q.wait()
    elif out.numel() > 0:
        dist.recv(out, src=src, group=group)
    return out


def gather_batch(shard, nbatch, dst=0, group=None):
    """Inverse of scatter_batch: returns the global tensor on `dst`, None elsewhere."""
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    if rk == dst:
        out = torch.empty((nbatch,) + tuple(shard.shape[1:]), dtype=shard.dtype, device=shard.device)
        for r in range(ws):
            lo, hi = shard_bounds(nbatch, ws, r)
            if r == dst:
                out[lo:hi].copy_(shard)
            elif hi > lo:
                dist.recv(out[lo:hi], src=r, group=group)
        return out
    if shard.numel() > 0:
        dist.send(shard.contiguous(), dst=dst, group=group)
    return None


def allreduce_shared_grad(local_mean, local_n, nbatch, group=None):
    """Global batch mean of a shared-input gradient from per-shard means (weights = shard sizes)."""
    g = local_mean * (float(local_n) / float(nbatch))
    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    return g


def sharded_qp(solve, Q, p, G, h, A, b, nbatch, src=0, group=None, device=None):
    """Scatter fully batched inputs from `src`, run `solve(Q,p,G,h,A,b) -> zhat` on every rank's shard,
    gather zhat on `src`. `solve` is `QPFunction(...)` on the GPUs; tests inject a CPU stand-in."""
    rk = dist.get_rank(group)
    parts = []
    for t in (Q, p, G, h, A, b):
        empty = [t is None or t.numel() == 0] if rk == src else [None]
        dist.broadcast_object_list(empty, src=src, group=group)
        parts.append(None if empty[0] else scatter_batch(t, nbatch, src, group, device))
    e = torch.Tensor()
    z = solve(*[x if x is not None else e for x in parts])
    return gather_batch(z.detach(), nbatch, dst=src, group=group)


def _scatter_even(out, full, src, group):
    """dist.scatter of equal contiguous shards (NCCL and gloo both implement it); `full` only on `src`."""
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    chunks = list(full.chunk(ws, dim=0)) if rk == src else None
    dist.scatter(out, scatter_list=chunks, src=src, group=group)


def _gather_even(shard, full, dst, group):
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    chunks = list(full.chunk(ws, dim=0)) if rk == dst else None
    dist.gather(shard, gather_list=chunks, dst=dst, group=group)


def sharded_qp_timed(f, glob, nbatch, nz, nineq, device, include_comm=True, src=0, group=None, dl=None):
    """BASELINE.json config 5 as one job: rank `src` holds the fully batched Q, p, G, h (dict `glob`, None elsewhere);
    they are scattered (NCCL scatter over NVLink; equal shards - ragged batches go through `sharded_qp`), every rank
    runs `f` forward + backward on its shard, z* is gathered on `src`; the per-sample gradients stay on the rank that
    produced them (gathering 1.3 GB of dQ/dG is the one thing a data-parallel consumer never needs).
    Returns dict(ms = device time of the region, max over ranks; z = gathered z* on `src`; grads = local gradients).
    include_comm=False times the solve alone (scatter before the first event, gather after the second)."""
    ws, rk = dist.get_world_size(group), dist.get_rank(group)
    assert nbatch % ws == 0, "sharded_qp_timed needs equal shards (use sharded_qp for ragged batches)"
    nloc = nbatch // ws
    f64 = dict(dtype=torch.float64, device=device)
    shapes = {"Q": (nz, nz), "p": (nz,), "G": (nineq, nz), "h": (nineq,)}
    loc = {k: torch.empty((nloc,) + shp, **f64) for k, shp in shapes.items()}
    zfull = torch.empty(nbatch, nz, **f64) if rk == src else None
    e = torch.empty(0, **f64)
    if dl is None:
        dl = torch.ones(nloc, nz, **f64)
    on_gpu = torch.device(device).type == "cuda"

    class _HostClock:       # gloo / CPU tests: same protocol as a CUDA event, wall clock
        def record(self):
            import time
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    ev0, ev1 = ((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if on_gpu
                else (_HostClock(), _HostClock()))

    def sync():
        if on_gpu:
            torch.cuda.synchronize()

    def scatter():
        for k in ("Q", "p", "G", "h"):
            _scatter_even(loc[k], glob[k] if rk == src else None, src, group)

    if include_comm:
        ev0.record()
        scatter()
    else:
        scatter()
        sync()
        ev0.record()
    t = {k: v.requires_grad_(True) for k, v in loc.items()}
    z = f(t["Q"], t["p"], t["G"], t["h"], e, e)
    z.backward(dl)
    zd = z.detach().contiguous()
    if include_comm:
        _gather_even(zd, zfull, src, group)
        ev1.record()
    else:
        ev1.record()
        _gather_even(zd, zfull, src, group)
    sync()
    ms = torch.tensor([ev0.elapsed_time(ev1)], **f64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
    return dict(ms=float(ms.item()), z=zfull, grads={k: v.grad for k, v in t.items()}, nloc=nloc)
