"""world_size-2 gloo tests of the batch-sharding host logic (no GPU): shard bounds, scatter/gather of ragged
shards, and the shard-size-weighted mean that keeps the reference's `.mean(0)` rule for shared inputs.
The solver is injected (the oracle), because the product has no CPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, nbatch, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qpth_b200 import parallel
        from qpth_b200.problems import random_qp_batch
        from oracle import pdipm_oracle as orc
        pr = random_qp_batch(nbatch, 6, 4, 2, seed=5) if rank == 0 else None
        T = {k: (torch.from_numpy(pr[k]) if rank == 0 else None) for k in ("Q", "p", "G", "h", "A", "b")}

        def solve(Q, p, G, h, A, b):
            r = orc.qp_solve(Q.numpy(), p.numpy(), G.numpy(), h.numpy(), A.numpy(), b.numpy(), per_qp=True)
            return torch.from_numpy(r["zhat"])

        z = parallel.sharded_qp(solve, T["Q"], T["p"], T["G"], T["h"], T["A"], T["b"], nbatch)
        # shard-size weighted mean of a "shared gradient"
        lo, hi = parallel.shard_bounds(nbatch, world, rank)
        full = torch.arange(nbatch, dtype=torch.float64).view(-1, 1) * torch.ones(1, 3, dtype=torch.float64)
        g = parallel.allreduce_shared_grad(full[lo:hi].mean(0) if hi > lo else torch.zeros(3, dtype=torch.float64),
                                           hi - lo, nbatch)
        if rank == 0:
            ref = orc.qp_solve(pr["Q"], pr["p"], pr["G"], pr["h"], pr["A"], pr["b"], per_qp=True)["zhat"]
            results["zerr"] = float(np.abs(z.numpy() - ref).max())
            results["gerr"] = float((g - full.mean(0)).abs().max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nbatch", [5, 8])
def test_sharded_qp_world2(nbatch):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), nbatch, results), nprocs=2, join=True)
    assert results["zerr"] < 1e-12 and results["gerr"] < 1e-12


def test_shard_bounds_cover_batch():
    from qpth_b200.parallel import shard_bounds
    for n in (0, 1, 7, 128, 8191):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def _worker_c5(rank, world, port, results):
    """The config-5 job function (scatter -> fwd+bwd on the shard -> gather z*) under gloo with an injected
    differentiable CPU solver: exercises exactly the collectives bench.py runs over NCCL."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qpth_b200 import parallel
        from qpth_b200.problems import random_qp_batch
        nbatch, nz, nineq = 8, 6, 4
        pr = random_qp_batch(nbatch, nz, nineq, 0, seed=3)
        glob = {k: torch.from_numpy(pr[k]) for k in ("Q", "p", "G", "h")} if rank == 0 else None

        def f(Q, p, G, h, A, b):          # stand-in with the layer's signature: unconstrained minimiser -Q^-1 p
            return -torch.linalg.solve(Q, p.unsqueeze(-1)).squeeze(-1) + 0.0 * (G.sum((1, 2)) + h.sum(1)).unsqueeze(-1)

        for comm in (True, False):
            out = parallel.sharded_qp_timed(f, glob, nbatch, nz, nineq, torch.device("cpu"), include_comm=comm)
            assert out["ms"] >= 0.0 and out["nloc"] == nbatch // world
            assert tuple(out["grads"]["Q"].shape) == (nbatch // world, nz, nz)
            if rank == 0:
                ref = -np.linalg.solve(pr["Q"], pr["p"][:, :, None])[:, :, 0]
                results["zerr_%d" % comm] = float(np.abs(out["z"].numpy() - ref).max())
            else:
                assert out["z"] is None
    finally:
        dist.destroy_process_group()


def test_c5_job_function_world2():
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker_c5, args=(2, _free_port(), results), nprocs=2, join=True)
    assert results["zerr_1"] < 1e-9 and results["zerr_0"] < 1e-9     # (Q is ill-conditioned: two LAPACK paths)
