"""N>1 host logic on CPU: column sharding and the per-step scalar allreduce (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from our_first_climate_model_b200 import distributed as rdist


def test_shard_ranges_partition_the_ensemble():
    for ncol in (1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [rdist.shard_range(ncol, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == ncol
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    rdist.init("gloo")
    ncol = 1001
    lo, hi = rdist.shard_range(ncol, rank, world)
    rng = np.random.default_rng(0)
    diag = rng.normal(size=(3, ncol, 4))  # [step][column][toa, dT, dE, dt] as the kernel writes them
    mine = diag[:, lo:hi]
    # what rcm_reduce_diag_kernel produces on each rank, then the collective
    sc = np.stack([mine[:, :, 0].sum(1), np.abs(mine[:, :, 1]).max(1), (np.abs(mine[:, :, 1]) < 0.5).sum(1),
                   np.abs(mine[:, :, 2]).max(1)], axis=1)
    t = torch.from_numpy(sc.copy())
    rdist.allreduce_step_scalars(t)
    full = np.stack([diag[:, :, 0].sum(1), np.abs(diag[:, :, 1]).max(1), (np.abs(diag[:, :, 1]) < 0.5).sum(1),
                     np.abs(diag[:, :, 2]).max(1)], axis=1)
    ok = np.allclose(t.numpy(), full, rtol=1e-12, atol=1e-12)
    # the asynchronous per-step exchange: same numbers, several steps in flight, ring reuse
    ex = rdist.StepScalarExchange(torch.device("cpu"), ring=2)
    tickets = [ex.submit(torch.from_numpy(sc[k].copy())) for k in range(3)]
    ok = ok and np.allclose(ex.result(tickets[2]).numpy(), full[2], rtol=1e-12, atol=1e-12)
    ok = ok and np.allclose(ex.result(tickets[1]).numpy(), full[1], rtol=1e-12, atol=1e-12)
    ok = ok and np.allclose(ex.latest().numpy(), full[2], rtol=1e-12, atol=1e-12)
    ex.drain()
    mx = rdist.max_over_ranks(10.0 + rank)
    means = rdist.global_means(t, ncol)
    rdist.barrier()
    out.put((rank, bool(ok), mx, means["converged_fraction"], float(full[-1, 2]) / ncol))


def test_scalar_allreduce_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, mx, frac, want in res:
        assert ok and mx == 11.0 and abs(frac - want) < 1e-12
