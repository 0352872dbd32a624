"""Cross-process parity of the sharded count against the oracle, every exchange mode (3 chunked exchange over the copy
engines, 2 sharded scatter, 1 two-pass fused route, 0 all-to-all).

  * test_two_processes_one_gpu_*: two rank processes on ONE device, CUDA-IPC peer buffers between them, gloo for the
    small collectives (NCCL refuses two ranks on one device).  Runs on the single-GPU test tier, so the cross-process
    visibility of the exchange (IPC mappings, peer copies, barriers) is covered there.
  * test_two_gpu_*: two devices over NCCL / NVLink.  Needs >= 2 B200s (`gpurun --gpus 2`); skipped otherwise."""
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
K, READS = 31, 200_000
MODES = (3, 2, 1, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _reads(rank, n=READS):
    from orion_kmer_b200 import synth
    g = synth.genome(90, 2_000_000)
    return synth.reads(g, 91, n, first_read=rank * READS, threads=4), synth.read_offsets(n)


def _worker(rank, world, port, ret, backend, one_device):
    import torch
    import torch.distributed as dist
    import orion_kmer_b200 as ok
    from orion_kmer_b200 import multi
    dev = 0 if one_device else rank
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}",
                                device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    ok.init(dev)
    # ranks of unequal batch size: the last one is a quarter short
    n = READS - (READS // 4 if rank == world - 1 else 0)
    bases, off = _reads(rank, n)
    d_b = torch.from_numpy(bases).cuda()
    d_o = torch.from_numpy(off.view(np.int64)).cuda()
    out, fallbacks = {}, {}
    for fused in MODES:
        sc = multi.ShardedCounter(ok, torch, dist, K, fused=fused, capacity_hint=0 if fused != 3 else 9_000_000)
        for _ in range(2):                      # twice: buffers are reused across steps
            sc.clear()
            sc.count_batch_device(d_b, len(bases), d_o, n)
        out[fused] = sc.counter.finish(1)
        fallbacks[fused] = sc.fallbacks
        sc.close()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((out, fallbacks), gathered, dst=0)
    if rank == 0:
        ret["tables"] = [g[0] for g in gathered]
        ret["fallbacks"] = [g[1] for g in gathered]
    dist.barrier()
    dist.destroy_process_group()


def _run(world, backend, one_device, oracle):
    import torch.multiprocessing as mp
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret, backend, one_device), nprocs=world, join=True)
        tables, fallbacks = ret["tables"], ret["fallbacks"]
    batches = [_reads(r, READS - (READS // 4 if r == world - 1 else 0)) for r in range(world)]
    all_bases = np.concatenate([b for b, _ in batches])
    all_off = np.arange(sum(len(o) - 1 for _, o in batches) + 1, dtype=np.uint64) * np.uint64(150)
    wk, wc = oracle.count_batch_mt(K, all_bases, all_off, 8)
    for fused in MODES:
        gk = np.concatenate([t[fused][0] for t in tables])
        gc = np.concatenate([t[fused][1] for t in tables])
        assert np.array_equal(gk, wk) and np.array_equal(gc, wc), f"fused={fused}"
        assert all(f[fused] == 0 for f in fallbacks), f"fused={fused} fell back to the all-to-all route: {fallbacks}"


def test_two_processes_one_gpu_all_exchange_modes_match_oracle(oracle):
    _run(2, "gloo", True, oracle)


def test_two_gpu_sharded_count_matches_oracle(oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "nccl", False, oracle)
