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
