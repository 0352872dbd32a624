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


# ------------------------- sets over several ranks: build, reshard, all-vs-all, query, classify --
class _HostKmerSet:
    """numpy stand-in for ok.KmerSet in the multi.py drivers of the sharded set algebra: a sorted duplicate-free
    array.  Set contents and per-read hits come from the oracle, the owner rule from the library's own host hook
    (okx_owner_of: the rule ok_set_shard_bounds applies on the device).  What the test checks is multi.py's
    bookkeeping: which set lives where, how the send buffer is packed, where each received slice starts, and that the
    per-shard integers add up to the whole."""
    K = 21

    def __init__(self, keys):
        self.keys = np.ascontiguousarray(keys, dtype=np.uint64)

    def __len__(self):
        return len(self.keys)

    def close(self):
        pass

    @classmethod
    def build_many(cls, k, batches):
        import oracle
        return [cls(oracle.kmer_set_batch(k, b, o)) for b, o in batches]

    @classmethod
    def from_sorted_device(cls, k, ptr, n):
        buf = (C.c_uint64 * max(n, 1)).from_address(ptr) if n else None
        keys = np.frombuffer(buf, dtype=np.uint64, count=n).copy() if n else np.zeros(0, np.uint64)
        assert n < 2 or np.all(keys[1:] > keys[:-1]), "a received slice must be strictly ascending"
        return cls(keys)

    def shard_bounds(self, n_ranks):
        import orion_kmer_b200 as ok
        owners = np.zeros(len(self.keys), dtype=np.int32)
        if len(self.keys):
            assert ok.lib().okx_owner_of(ok._ptr(self.keys), len(self.keys), self.K, n_ranks, ok._ptr(owners)) == 0
            assert np.all(np.diff(owners) >= 0), "owners must be monotone in the key"
        return np.searchsorted(owners, np.arange(n_ranks + 1), "left").astype(np.uint64)

    def copy_keys_device(self, first, n, ptr):
        C.memmove(ptr, self.keys[first:first + n].ctypes.data, 8 * n)

    def probe_reads(self, bases, off, norm_mode):
        import oracle
        return oracle.query_hits(self.keys, self.K, bases, off, 1).astype(np.uint32)


class _HostOk:
    KmerSet = _HostKmerSet
    finish_all_vs_all = staticmethod(_HostSets.finish_all_vs_all)

    @staticmethod
    def all_vs_all_part(sets, part, n_parts):
        return _HostSets.all_vs_all_part([s.keys for s in sets], part, n_parts)

    @staticmethod
    def probe_counts_many(shards, kmers, counts):
        hit = [np.isin(kmers, s.keys, assume_unique=True) for s in shards]
        return (np.array([h.sum() for h in hit], np.uint64), np.array([counts[h].sum() for h in hit], np.uint64))


N_GENOMES = 7


def _genome_batch(i):
    from orion_kmer_b200 import synth
    base = synth.genome(90, 30_000)
    g = base if i == 0 else (synth.mutate(base, 90 + i, 200 * i) if i < 5 else synth.genome(95 + i, 12_000 + 1000 * i))
    if i == 6:
        g = g[:15]                                  # shorter than k: an empty set
    return np.ascontiguousarray(g, dtype=np.uint8), np.array([0, len(g) // 3, len(g)], np.uint64)


def _query_reads():
    from orion_kmer_b200 import synth
    return (np.concatenate([synth.reads(_genome_batch(0)[0], 5, 200, threads=1), synth.reads(synth.genome(77, 30_000), 6, 100, threads=1)]),
            synth.read_offsets(300))


def _sharded_sets_worker(rank, world, port, ret):
    import oracle
    from orion_kmer_b200 import multi
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method=f"tcp://127.0.0.1:{port}")
    try:
        k = _HostKmerSet.K
        mine = multi.build_sets(_HostOk, dist, k, N_GENOMES, _genome_batch)
        assert sorted(mine) == list(range(rank, N_GENOMES, world))          # genome i on rank i mod N, no exchange
        sizes, full = multi.all_vs_all_sharded(_HostOk, torch, dist, k, mine, N_GENOMES, key_device="cpu")
        shards, sizes2 = multi.reshard_sets(_HostOk, torch, dist, k, mine, N_GENOMES, key_device="cpu")
        assert np.array_equal(sizes, sizes2)
        # every shard holds keys this rank owns, and nothing else
        for s in shards:
            b = s.shard_bounds(world)
            assert int(b[rank]) == 0 and int(b[rank + 1]) == len(s)
        union = _HostKmerSet(oracle.set_union([s.keys for s in shards]))
        bases, off = _query_reads()
        hits = multi.query_sharded(_HostOk, torch, dist, union, bases, off)
        wk, wc = oracle.count_batch(k, bases, off)
        matched, depth, ref_sizes = multi.classify_sharded(_HostOk, torch, dist, shards, wk, wc)
        per_rank = [None] * world
        dist.all_gather_object(per_rank, [len(s) for s in shards])
        if rank == 0:
            ret.update(sizes=sizes, full=full, hits=hits, matched=matched, depth=depth, ref_sizes=ref_sizes, per_rank=per_rank)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_set_algebra_adds_up_to_the_whole(oracle, world):
    """SURVEY 8e rows 2-4 on CPU (gloo): genome-per-rank build, identical key-range sharding of every set through one
    all-to-all, then all-vs-all / query / classify as sums over the shards -- against the oracle on the whole sets."""
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_sharded_sets_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        got = dict(ret)
    k = _HostKmerSet.K
    sets = [oracle.kmer_set_batch(k, *_genome_batch(i)) for i in range(N_GENOMES)]
    assert len(sets[6]) == 0 and len(sets[0]) > 20_000
    assert np.array_equal(got["sizes"], np.array([len(s) for s in sets], np.uint64))
    for i in range(N_GENOMES):
        for j in range(N_GENOMES):
            want = len(sets[i]) if i == j else len(np.intersect1d(sets[i], sets[j], assume_unique=True))
            assert got["full"][i, j] == want, (i, j)
    bases, off = _query_reads()
    assert np.array_equal(got["hits"].astype(np.uint64), oracle.query_hits(oracle.set_union(sets), k, bases, off, 2))
    assert got["hits"][:200].min() > 40 and got["hits"][200:].max() < 5          # reads of genome 0 hit, unrelated ones do not
    wk, wc = oracle.count_batch(k, bases, off)
    for i, s in enumerate(sets):
        m, d = oracle.classify_ref(wk, wc, s)                                      # classify.rs:224-236 on the whole reference
        assert (got["matched"][i], got["depth"][i], got["ref_sizes"][i]) == (m, d, len(s)), i
    # the ranks' shards of a set partition it
    assert [sum(r[i] for r in got["per_rank"]) for i in range(N_GENOMES)] == [len(s) for s in sets]
    shares = np.array([sum(r) for r in got["per_rank"]], dtype=float)
    assert shares.max() / shares.mean() < 1.25, shares                             # owner ranges balanced by the canonical prior
