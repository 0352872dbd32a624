"""N>1 host logic on CPU: world_size-2 gloo run of the owner rule + exchange + concatenation.
The k-mers come from the host emulation of the extraction kernel (same __host__ __device__ code),
the owner of each k-mer from the same routine k_route uses; local counting is numpy.  What is
checked: owners form contiguous key ranges, the exchange delivers every k-mer to its owner, and the
ranks' sorted tables concatenate to the oracle's global table."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

K, READS = 21, 3000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_reads(rank):
    from orion_kmer_b200 import synth
    g = synth.genome(7, 60_000)
    return synth.reads(g, 8, READS, first_read=rank * READS, threads=1), synth.read_offsets(READS)


def _worker(rank, world, port, ret):
    import orion_kmer_b200 as ok
    from orion_kmer_b200 import multi
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    try:
        bases, off = _rank_reads(rank)
        keys = np.zeros(len(bases), dtype=np.uint64)
        n = C.c_uint64()
        assert ok.lib().okx_emulate_extract(ok._ptr(bases), len(bases), ok._ptr(off), READS, K, ok.NORMALIZED,
                                            ok._ptr(keys), len(keys), C.byref(n)) == 0
        keys = keys[:n.value]
        owners = np.zeros(len(keys), dtype=np.int32)
        assert ok.lib().okx_owner_of(ok._ptr(keys), len(keys), K, world, ok._ptr(owners)) == 0
        order = np.argsort(owners, kind="stable")
        send = torch.from_numpy(keys[order].view(np.int64).copy())
        counts = np.bincount(owners, minlength=world)
        recv, recv_counts = multi.exchange(dist, torch, send, counts)
        mine = recv.numpy().view(np.uint64)
        chk = np.zeros(len(mine), dtype=np.int32)
        assert ok.lib().okx_owner_of(ok._ptr(mine), len(mine), K, world, ok._ptr(chk)) == 0
        assert np.all(chk == rank), "a k-mer reached a rank that does not own it"
        uk, uc = np.unique(mine, return_counts=True)
        out = [None] * world if rank == 0 else None
        dist.gather_object((uk, uc.astype(np.uint64), int(len(keys)), int(sum(recv_counts))), out, dst=0)
        if rank == 0:
            ret["tables"] = out
    finally:
        dist.destroy_process_group()


def test_two_rank_exchange_concatenates_to_global_table(oracle):
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        tables = ret["tables"]
    all_bases = np.concatenate([_rank_reads(r)[0] for r in range(world)])
    all_off = np.arange(world * READS + 1, dtype=np.uint64) * np.uint64(150)
    wk, wc = oracle.count_batch(K, all_bases, all_off)
    gk = np.concatenate([t[0] for t in tables])
    gc = np.concatenate([t[1] for t in tables])
    assert np.array_equal(gk, wk) and np.array_equal(gc, wc)          # rank order == key order
    assert sum(t[2] for t in tables) == sum(t[3] for t in tables) == int(wc.sum())
    # both ranks hold a comparable share (owner ranges are balanced by the canonical prior)
    shares = [len(t[0]) / len(wk) for t in tables]
    assert all(0.4 < s < 0.6 for s in shares), shares


# ------------------------------------------------------- all-vs-all compare across ranks --
class _HostSets:
    """numpy stand-in for the device side of multi.all_vs_all: the same pair dealing as ok_sets_all_vs_all_part
    (pairs i < j numbered row by row, pair p belongs to part p % n_parts), intersections by np.intersect1d"""

    @staticmethod
    def all_vs_all_part(sets, part, n_parts):
        n = len(sets)
        sizes = np.array([len(s) for s in sets], dtype=np.uint64)
        inter = np.zeros((n, n), dtype=np.uint64)
        p = 0
        for i in range(n):
            for j in range(i + 1, n):
                if p % n_parts == part:
                    inter[i, j] = len(np.intersect1d(sets[i], sets[j], assume_unique=True))
                p += 1
        return sizes, inter

    @staticmethod
    def finish_all_vs_all(sizes, upper):
        import orion_kmer_b200 as ok
        return ok.finish_all_vs_all(sizes, upper)


def _sets():
    rng = np.random.default_rng(12)
    pool = np.unique(rng.integers(0, 2 ** 42, 40_000, dtype=np.uint64))
    return [pool[rng.random(len(pool)) < f] for f in (0.5, 0.3, 0.7, 0.0, 0.5, 1.0, 0.1)]


def _ava_worker(rank, world, port, ret):
    from orion_kmer_b200 import multi
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    try:
        sizes, full = multi.all_vs_all(_HostSets, torch, dist, _sets())
        if rank == 0:
            ret["sizes"], ret["full"] = sizes, full
    finally:
        dist.destroy_process_group()


def test_two_rank_all_vs_all_matches_single_process():
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_ava_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        sizes, full = ret["sizes"], ret["full"]
    sets = _sets()
    n = len(sets)
    for i in range(n):
        assert sizes[i] == len(sets[i]) == full[i, i]
        for j in range(n):
            assert full[i, j] == len(np.intersect1d(sets[i], sets[j], assume_unique=True)), (i, j)


def _coll_worker(rank, world, port):
    from orion_kmer_b200 import multi
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    try:
        c = multi.Coll(dist, torch)
        t = torch.tensor([rank + 1], dtype=torch.int32)
        assert c.all_reduce(t, dist.ReduceOp.MIN).item() == 1
        inp = torch.arange(4, dtype=torch.int32) + 10 * rank
        out = torch.empty(8, dtype=torch.int32)
        assert c.all_gather(out, inp).tolist() == [0, 1, 2, 3, 10, 11, 12, 13]
        mine = torch.empty(2, dtype=torch.int32)
        assert c.reduce_scatter(mine, inp.clone()).tolist() == ([10, 12] if rank == 0 else [14, 16])
        c.barrier()
    finally:
        dist.destroy_process_group()


def test_collective_helper_without_nccl():
    """multi.Coll stages the exchange's small collectives through host memory on a backend without device support
    (gloo: what the two-processes-on-one-GPU test uses); same results as the NCCL forms."""
    mp.spawn(_coll_worker, args=(2, _free_port()), nprocs=2, join=True)


def _flags_worker(rank, world, port, ret):
    from orion_kmer_b200 import multi
    import time
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    try:
        f = multi.HostFlags(dist, multi.Coll(dist, torch))
        assert f.slots is not None, "shared memory unavailable"
        order = []
        for i in range(50):
            if rank == i % world:
                time.sleep(0.002)               # the late rank: the others must wait for it
            order.append(time.perf_counter())
            f.barrier()
        t_after = time.perf_counter()
        out = [None] * world
        dist.all_gather_object(out, (order, t_after))
        # nobody left barrier i before everybody had entered it
        for i in range(49):
            assert max(o[0][i] for o in out) <= min(o[0][i + 1] for o in out) + 1e-3
        f.close()
        if rank == 0:
            ret["ok"] = True
    finally:
        dist.destroy_process_group()


def test_host_flags_barrier():
    """the shared-memory barrier the chunked exchange paces its receive pipeline with"""
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_flags_worker, args=(3, _free_port(), ret), nprocs=3, join=True)
        assert ret.get("ok")
