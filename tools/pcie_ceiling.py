#!/usr/bin/env python
"""What the host side allows: N ranks (one per GPU) concurrently copy 1.58 GB in and 3.44 GB out of page-locked memory
-- the bytes one rank of the count job moves per step (bench.py e2e).  Run under torchrun like bench.py:
  python -m torch.distributed.run --nproc-per-node N tools/pcie_ceiling.py
Prints one JSON line (rank 0): GB/s per rank and aggregate for H2D alone, D2H alone, both at once, and the time the
e2e copies of one step would take at those rates."""
import json
import os
import time

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
IN, OUT = 1_580_000_008, 3_440_474_624
h_in = torch.empty(IN, dtype=torch.uint8).pin_memory()
h_out = torch.empty(OUT, dtype=torch.uint8).pin_memory()
d_in = torch.empty(IN, dtype=torch.uint8, device="cuda")
d_out = torch.empty(OUT, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(what, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if what in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if what in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best


res = {}
for what, nbytes in (("h2d", IN), ("d2h", OUT), ("both", IN + OUT)):
    dt = run(what)
    res[what] = {"seconds": dt, "GBs_per_rank": nbytes / dt / 1e9, "GBs_aggregate": nbytes * world / dt / 1e9}
if int(os.environ.get("RANK", "0")) == 0:
    res["n_gpus"] = world
    res["e2e_copy_floor_ms_sequential"] = (res["h2d"]["seconds"] + res["d2h"]["seconds"]) * 1e3
    res["e2e_copy_floor_ms_full_duplex"] = res["both"]["seconds"] * 1e3
    print(json.dumps(res))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
