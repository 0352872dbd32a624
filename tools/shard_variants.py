#!/usr/bin/env python
"""Multi-GPU tuning run: the sharded count step under several geometries, one process group, data generated once.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/shard_variants.py [--reads R] [--steps K]

Variants: capacity hint, ORION_SHARD_B1 (level-1 bits of the sender), ORION_SUB_BITS (sub-partition bits).  Prints one line per variant on rank 0 (device-resident step, max over ranks).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    import orion_kmer_b200 as ok
    from orion_kmer_b200 import multi, synth
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ok.init(local)
    n_reads = args.reads
    g, bases, off = bench.make_workload(ok, synth, n_reads, n_reads * 5 * world, first_read=rank * n_reads)
    n_bases = len(bases)
    d_bases = torch.from_numpy(bases).cuda()
    d_off = torch.from_numpy(off.view(np.int64)).cuda()
    hint = int(n_bases * 0.17)
    # (name, capacity hint, environment of the geometry)
    variants = [("tma push", hint, {"ORION_PUSH_TMA": "1"}), ("register push", hint, {"ORION_PUSH_TMA": "0"}),
                ("tma bits15", hint, {"ORION_PUSH_TMA": "1", "ORION_SUB_BITS": "15"}),
                ("tma b1=6", hint, {"ORION_PUSH_TMA": "1", "ORION_SHARD_B1": "6"}),
                ("tma no hint", 0, {"ORION_PUSH_TMA": "1"})]
    for name, h, env in variants:
        for k in ("ORION_SHARD_B1", "ORION_SUB_BITS", "ORION_PUSH_TMA"):
            os.environ.pop(k, None)
        os.environ.update(env)          # read by ok_shard_geometry when the first batch is counted
        sc = multi.ShardedCounter(ok, torch, dist, 31, fused=2, capacity_hint=h)
        acc = []
        for i in range(args.warmup + args.steps):
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            sc.clear()
            sc.count_batch_device(d_bases, n_bases, d_off, n_reads)
            sc.counter.finish_device(1)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if i >= args.warmup:
                st = sc.counter.stats()
                acc.append((float(dt.item()) * 1e3, st["ms_scatter1"], st["ms_scatter2"], st["ms_count"], st["ms_compact"], st["ms_push"]))
        m = np.mean(np.array(acc), axis=0)
        if rank == 0:
            print(f"N={world} {name:10s} geom={sc.geom['sub_bits']}/{sc.geom['l1_bits']} step {m[0]:.2f} ms  "
                  f"{world * n_bases / m[0] / 1e6:.1f} G bases/s  scatter+push {m[1]:.2f} level2 {m[2]:.2f} count {m[3]:.2f} "
                  f"compact {m[4]:.2f} push {m[5]:.2f} fallbacks {getattr(sc, 'fallbacks', 0)}", flush=True)
        sc.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
