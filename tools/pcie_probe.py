#!/usr/bin/env python
"""Host<->device copy bandwidth per rank, one rank at a time and all ranks at once (torchrun, one rank per GPU).
Explains the spread of bench.py's e2e over the ranks of one box: rcm_step_host moves 11 MB up and 43.5 MB down per step.
Usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (bind_to_gpu_numa: the placement the bench uses)


def bw(dst, src, reps):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    return dst.numel() * 8 * reps / (s.elapsed_time(e) * 1e-3) / 1e9


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    numa = bench.bind_to_gpu_numa(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_down, n_up = 65536 * 83, 65536 * 21
    h_down, h_up = torch.empty(n_down, dtype=torch.float64).pin_memory(), torch.empty(n_up, dtype=torch.float64).pin_memory()
    d_down, d_up = torch.zeros(n_down, dtype=torch.float64, device="cuda"), torch.zeros(n_up, dtype=torch.float64, device="cuda")
    out = {"rank": rank, "numa": numa}
    for r in range(world):  # one rank at a time
        if world > 1:
            dist.barrier()
        if r == rank:
            out["alone_d2h"], out["alone_h2d"] = bw(h_down, d_down, 40), bw(d_up, h_up, 40)
    if world > 1:
        dist.barrier()
    out["all_d2h"] = bw(h_down, d_down, 80)
    if world > 1:
        dist.barrier()
    out["all_h2d"] = bw(d_up, h_up, 80)
    rows = [None] * world
    if world > 1:
        dist.all_gather_object(rows, out)
    else:
        rows = [out]
    if rank == 0:
        for row in rows:
            print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in row.items()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
