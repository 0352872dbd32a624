"""2-GPU parity of the sharded count (all three exchange modes: sharded scatter, two-pass fused route, NCCL all-to-all) against the oracle.  Needs >= 2 B200s:
run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`; skipped otherwise."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
K, READS = 31, 200_000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _reads(rank):
    from orion_kmer_b200 import synth
    g = synth.genome(90, 2_000_000)
    return synth.reads(g, 91, READS, first_read=rank * READS, threads=4), synth.read_offsets(READS)


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    import orion_kmer_b200 as ok
    from orion_kmer_b200 import multi
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}",
                            device_id=torch.device("cuda", rank))
    ok.init(rank)
    bases, off = _reads(rank)
    d_b = torch.from_numpy(bases).cuda()
    d_o = torch.from_numpy(off.view(np.int64)).cuda()
    out = {}
    for fused in (2, 1, 0):
        sc = multi.ShardedCounter(ok, torch, dist, K, fused=fused)
        for _ in range(2):                      # twice: buffers are reused across steps
            sc.clear()
            sc.count_batch_device(d_b, len(bases), d_o, READS)
        out[fused] = sc.counter.finish(1)
        sc.close()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(out, gathered, dst=0)
    if rank == 0:
        ret["tables"] = gathered
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_count_matches_oracle(oracle):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        tables = ret["tables"]
    all_bases = np.concatenate([_reads(r)[0] for r in range(world)])
    all_off = np.arange(world * READS + 1, dtype=np.uint64) * np.uint64(150)
    wk, wc = oracle.count_batch(K, all_bases, all_off)
    for fused in (2, 1, 0):
        gk = np.concatenate([t[fused][0] for t in tables])
        gc = np.concatenate([t[fused][1] for t in tables])
        assert np.array_equal(gk, wk) and np.array_equal(gc, wc), f"fused={fused}"
