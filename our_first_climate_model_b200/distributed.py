"""Multi-GPU plumbing: one process per GPU, columns sharded, ONE tiny collective per step.

Columns of the ensemble are independent (no horizontal coupling anywhere in the reference's
main.cpp), so the data path has no collective at all.  The only traffic is the allreduce of the
four per-step ensemble scalars (rcm_step_scalars): sums for the TOA imbalance and the converged
count, maxima for the temperature change and |dE|.  torch.distributed supplies the transport
(NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

SUM_IDX = (0, 2)  # toa_net_sum, n_converged
MAX_IDX = (1, 3)  # max_dT, max_abs_dE


def init(backend: str = "nccl"):
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group(backend=backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def barrier():
    if dist.is_initialized():
        dist.barrier()


def shard_range(ncol: int, rank: int, world: int):
    """Contiguous block of columns owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(ncol, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DevPtr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def device_view(ptr: int, n: int) -> torch.Tensor:
    """Zero-copy float64 view of `n` doubles at a device address (the solver's scalar buffer)."""
    return torch.as_tensor(_DevPtr(ptr, n), device="cuda")


def allreduce_step_scalars(t: torch.Tensor) -> torch.Tensor:
    """In-place allreduce of a [..., 4] tensor laid out as rcm_step_scalars."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t
    v = t.view(-1, 4)
    s = v[:, list(SUM_IDX)].contiguous()
    m = v[:, list(MAX_IDX)].contiguous()
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    v[:, list(SUM_IDX)] = s
    v[:, list(MAX_IDX)] = m
    return t


def require_shared_stream(solver) -> None:
    """The scalars are read by torch / NCCL on torch's current stream: the solver must queue its kernels on that very
    stream (rcm_set_stream), otherwise the copy races the reduce kernel and stale scalars travel."""
    if solver is None or not torch.cuda.is_available():
        return
    if getattr(solver, "stream_ptr", None) != torch.cuda.current_stream().cuda_stream:
        raise RuntimeError("call solver.set_stream(s.cuda_stream) with a torch.cuda.Stream s that is torch's current "
                           "stream (the legacy default stream cannot be shared with the solver)")


class StepScalarExchange:
    """The per-step collective without a per-step stall.

    Every step, `submit(view)` copies the rank's four scalars into a ring slot and issues ONE asynchronous
    all_gather of that slot (32 bytes per rank over NVLink); NCCL runs it on its own stream behind the work already
    queued, so the next step's kernel starts immediately.  Sums and maxima over the ranks are taken from the gathered
    rows when somebody looks (`result(k)`, `latest()`): the decision "is the ensemble stationary" needs them every
    few hundred steps, not every step.  A slot is reused only after its collective has completed.
    With one rank (or no process group) it just keeps references to the local scalars.
    `solver`: when given, every submit() checks that the solver launches on torch's current stream."""

    def __init__(self, device, ring: int = 32, solver=None):
        self.solver = solver
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.ring = ring
        self.send = torch.zeros(ring, 4, dtype=torch.float64, device=device)
        self.recv = torch.zeros(ring, self.world, 4, dtype=torch.float64, device=device)
        self.work = [None] * ring
        self.count = 0

    def submit(self, scalars: torch.Tensor) -> int:
        """scalars: float64[4] (layout of rcm_step_scalars) of the step just queued. Returns the step's ticket."""
        require_shared_stream(self.solver)
        k = self.count % self.ring
        if self.work[k] is not None:
            self.work[k].wait()
            self.work[k] = None
        if self.world > 1:
            self.send[k].copy_(scalars.view(-1)[-4:])
            self.work[k] = dist.all_gather_into_tensor(self.recv[k].view(-1), self.send[k], async_op=True)
        else:
            self.recv[k, 0].copy_(scalars.view(-1)[-4:])  # one rank: the slot is the rank's own scalars
        self.count += 1
        return self.count - 1

    def result(self, ticket: int) -> torch.Tensor:
        """Reduced scalars [toa_net_sum, max_dT, n_converged, max_abs_dE] of a submitted step (must still be in the ring)."""
        assert self.count - self.ring <= ticket < self.count, "ticket has left the ring"
        k = ticket % self.ring
        if self.work[k] is not None:
            self.work[k].wait()
            self.work[k] = None
        rows = self.recv[k]
        out = rows.sum(dim=0)
        mx = rows.max(dim=0).values
        out[list(MAX_IDX)] = mx[list(MAX_IDX)]
        return out

    def latest(self) -> torch.Tensor:
        return self.result(self.count - 1)

    def drain(self):
        for k, w in enumerate(self.work):
            if w is not None:
                w.wait()
                self.work[k] = None


def max_over_ranks(x: float) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def global_means(scalars: torch.Tensor, ncol_total: int) -> dict:
    """Ensemble means after the allreduce: mean TOA imbalance, converged fraction, max changes."""
    v = scalars.view(-1, 4)[-1]
    return {"toa_net_mean": float(v[0]) / ncol_total, "max_dT": float(v[1]),
            "converged_fraction": float(v[2]) / ncol_total, "max_abs_dE": float(v[3])}


def run_to_equilibrium(solver, ncol_total: int, max_steps: int, check_every: int = 100) -> dict:
    """The RCE driver loop over all ranks: every rank advances its own shard in blocks of `check_every` fused steps
    (one launch per block, nothing crosses GPUs meanwhile); after each block ONE allreduce of the block's scalars
    decides - identically on every rank - whether the whole ensemble is stationary (n_converged == ncol_total).
    With one rank this is rcm_run_to_equilibrium."""
    require_shared_stream(solver)
    done = 0
    means = None
    while done < max_steps:
        n = min(check_every, max_steps - done)
        ptr = solver.advance_async(n)
        sc = device_view(ptr, 4 * n)
        allreduce_step_scalars(sc)
        done += n
        means = global_means(sc, ncol_total)  # .item() inside: waits for the block
        if means["converged_fraction"] >= 1.0:
            break
    return {"steps": done, **(means or {})}
