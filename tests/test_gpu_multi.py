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


# ---------------------------------------------------------------- multi-GPU set algebra --
N_GENOMES, SET_K = 7, 21


def _genome(i):
    from orion_kmer_b200 import synth
    base = synth.genome(200 + i % 3, 300_000 + 10_000 * (i % 3))
    return base if i < 3 else synth.mutate(base, i, 5_000 * i)


def _probe_reads():
    from orion_kmer_b200 import synth
    n = 4000
    return np.concatenate([synth.reads(_genome(0), 210, n), synth.reads(synth.genome(299, 200_000), 211, n)]), synth.read_offsets(2 * n)


def _set_worker(rank, world, port, ret, backend, one_device):
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
    # e-2: genome-per-GPU build, no exchange
    mine = multi.build_sets(ok, dist, SET_K, N_GENOMES, lambda i: (_genome(i), np.array([0, len(_genome(i))], np.uint64)))
    assert sorted(mine) == list(range(rank, N_GENOMES, world))
    # e-3: identical key-range sharding of every set, full matrix per shard, one all-reduce
    sizes, inter = multi.all_vs_all_sharded(ok, torch, dist, SET_K, mine, N_GENOMES)
    # e-4: the database sharded by key range, reads replicated, all-reduce of the integer results
    shards, sizes2 = multi.reshard_sets(ok, torch, dist, SET_K, mine, N_GENOMES)
    union = ok.KmerSet.union(shards)
    reads, off = _probe_reads()
    hits = multi.query_sharded(ok, torch, dist, union, reads, off, ok.RAW)
    pk, pc = ok.count_fastx(SET_K, [b">q\n" + bytes(reads[:150 * 2000])])
    matched, depth, ref_sizes = multi.classify_sharded(ok, torch, dist, shards, pk, pc)
    if rank == 0:
        ret["sets"] = dict(sizes=sizes, inter=inter, sizes2=sizes2, hits=hits, pk=pk, pc=pc, matched=matched, depth=depth,
                           ref_sizes=ref_sizes)
    dist.barrier()
    dist.destroy_process_group()


def _run_sets(world, backend, one_device, oracle):
    import torch.multiprocessing as mp
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_set_worker, args=(world, _free_port(), ret, backend, one_device), nprocs=world, join=True)
        got = ret["sets"]
    osets = [oracle.kmer_set_batch(SET_K, _genome(i), np.array([0, len(_genome(i))], np.uint64)) for i in range(N_GENOMES)]
    assert list(got["sizes"]) == [len(o) for o in osets] == list(got["sizes2"]) == list(got["ref_sizes"])
    for i in range(N_GENOMES):
        for j in range(N_GENOMES):
            assert got["inter"][i, j] == oracle.compare(osets[i], osets[j])["intersection_size"], (i, j)
    reads, off = _probe_reads()
    assert np.array_equal(got["hits"].astype(np.uint64), oracle.query_hits(oracle.set_union(osets), SET_K, reads, off, 8))
    for i, o in enumerate(osets):
        assert (int(got["matched"][i]), int(got["depth"][i])) == oracle.classify_ref(got["pk"], got["pc"], o), i


def test_two_processes_one_gpu_set_build_compare_query_classify(oracle):
    _run_sets(2, "gloo", True, oracle)


def test_two_gpu_set_build_compare_query_classify(oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_sets(2, "nccl", False, oracle)


def test_two_processes_one_gpu_all_exchange_modes_match_oracle(oracle):
    _run(2, "gloo", True, oracle)


def test_two_gpu_sharded_count_matches_oracle(oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, "nccl", False, oracle)
