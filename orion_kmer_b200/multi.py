"""Multi-GPU count: one process per GPU, k-mers exchanged by owner key range.

Per step and rank:  route (extract + bucket by owner, on the device)  ->  all-to-all of the 8-byte
k-mers over NCCL/NVLink  ->  count what arrived on the rank's own shard.  Owners are contiguous
ranges of the monotone k-mer position, so rank r's sorted output is the r-th slice of the global
sorted table and the ranks' outputs simply concatenate.

torch.distributed is the plumbing (process group, all_to_all_single); every kernel is ours.
`exchange` works on any backend (the CPU tests drive it over gloo).
"""
import json
import os
import time

import numpy as np


class Coll:
    """The handful of small collectives of the exchange.  Under NCCL they run on device tensors; under any other
    backend (gloo: the CPU tests, and the two-processes-on-ONE-GPU test that the single-GPU test tier can run) device
    tensors are staged through host memory.  Results are identical; only the plumbing differs."""

    def __init__(self, dist, torch, group=None):
        self.dist, self.torch, self.group = dist, torch, group
        self.native = dist.get_backend(group) == "nccl"

    def all_reduce(self, t, op=None):
        op = op if op is not None else self.dist.ReduceOp.SUM
        if self.native or not t.is_cuda:
            self.dist.all_reduce(t, op=op, group=self.group)
        else:
            h = t.cpu()
            self.dist.all_reduce(h, op=op, group=self.group)
            t.copy_(h)
        return t

    def all_gather(self, out, inp):
        """out: world * len(inp) elements, rank-major"""
        if self.native:
            self.dist.all_gather_into_tensor(out, inp, group=self.group)
            return out
        h = inp.cpu().contiguous()
        parts = [self.torch.empty_like(h) for _ in range(self.dist.get_world_size(self.group))]
        self.dist.all_gather(parts, h, group=self.group)
        out.copy_(self.torch.cat(parts))
        return out

    def reduce_scatter(self, out, inp):
        """out = this rank's slice of the element-wise sum of every rank's inp"""
        if self.native:
            self.dist.reduce_scatter_tensor(out, inp, group=self.group)
            return out
        h = inp.cpu()
        self.dist.all_reduce(h, group=self.group)
        r, n = self.dist.get_rank(self.group), out.numel()
        out.copy_(h[r * n:(r + 1) * n])
        return out

    def barrier(self):
        self.dist.barrier(group=self.group)


class HostFlags:
    """Host-side barrier of the ranks of one node in shared memory (a few microseconds; a collective through sockets
    or through the GPU costs 0.1-1 ms and, on the GPU, an SM slot).  Rank r bumps slot r; a barrier is over once every
    slot has reached the caller's tick.  Falls back to the process group's barrier when shared memory is unavailable."""

    def __init__(self, dist, coll):
        self.dist, self.coll = dist, coll
        self.rank, self.world, self.tick = dist.get_rank(), dist.get_world_size(), 0
        self.shm, self.slots = None, None
        try:
            from multiprocessing import shared_memory
            name = [None]
            if self.rank == 0:
                self.shm = shared_memory.SharedMemory(create=True, size=64 * self.world)
                self.shm.buf[:64 * self.world] = bytes(64 * self.world)
                name[0] = self.shm.name
            dist.broadcast_object_list(name, src=0)
            if self.rank != 0 and name[0]:
                self.shm = shared_memory.SharedMemory(name=name[0])
            ok = self.shm is not None
        except Exception:
            ok = False
        flags = [None] * self.world
        dist.all_gather_object(flags, ok)
        if all(flags):
            self.slots = np.ndarray((self.world, 8), dtype=np.int64, buffer=self.shm.buf)[:, 0]     # one cache line per rank
        coll.barrier()

    def barrier(self, timeout=60.0):
        self.tick += 1
        if self.slots is None:
            return self.coll.barrier()
        self.slots[self.rank] = self.tick
        t0, spins = time.perf_counter(), 0
        while int(self.slots.min()) < self.tick:
            spins += 1
            if spins % 4096 == 0 and time.perf_counter() - t0 > timeout:
                raise RuntimeError(f"rank {self.rank}: host barrier timed out (a peer rank is gone?)")

    def close(self):
        if self.shm is not None:
            self.slots = None
            try:
                self.coll.barrier()
                self.shm.close()
                if self.rank == 0:
                    self.shm.unlink()
            except Exception:
                pass
            self.shm = None


def exchange(dist, torch, send, send_counts, group=None):
    """send: 1-D int64 tensor holding the k-mers for rank 0, then rank 1, ... (send_counts each).
    -> (recv tensor, recv_counts list)."""
    world = dist.get_world_size(group)
    staged = send.is_cuda and dist.get_backend(group) != "nccl"
    dev = send.device
    if staged:
        send = send.cpu()
    sc = torch.as_tensor(np.asarray(send_counts, dtype=np.int64), device=send.device)
    rc = torch.empty(world, dtype=torch.int64, device=send.device)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.cpu().tolist()]
    recv = torch.empty(sum(recv_counts), dtype=torch.int64, device=send.device)
    dist.all_to_all_single(recv, send[:int(sum(send_counts))], output_split_sizes=recv_counts,
                           input_split_sizes=[int(x) for x in send_counts], group=group)
    return (recv.to(dev) if staged else recv), recv_counts


class ShardedCounter:
    """KmerCounter sharded over the ranks of a process group.

    fused=True (default): the routing kernel writes each owner's k-mers straight into that owner's
    receive buffer (CUDA-IPC peer memory over NVLink); the only collectives left are a world x world
    count matrix and a barrier.  fused=False: bucket locally, NCCL all_to_all_single, then count."""

    def __init__(self, ok, torch, dist, k, norm_mode=0, fused=3, capacity_hint=0):
        self.ok, self.torch, self.dist = ok, torch, dist
        self.coll = Coll(dist, torch)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        # capacity_hint = expected distinct k-mers of THIS rank's shard (0: none): sizes the sub-partitions for their
        # distinct keys (fewer bins per scatter level); a hint that proves too low costs one recount, then it is ignored
        self.counter = ok.KmerCounter(k, norm_mode, capacity_hint)
        self.hinted = capacity_hint > 0
        self.counter.set_shard(self.rank, self.world)
        # 3: chunked exchange (sample, then the extraction scatters chunk by chunk into per-chunk sub-blocks that plain
        #    copy-engine peer copies move under the next chunk's extraction); 2: sharded scatter (a copy warp in the
        #    extraction kernel pushes with SM stores); 1: two-pass fused route; 0: NCCL all-to-all
        self.fused = int(fused)
        self.fallbacks = 0
        self.margin = 1.25      # what a rank may receive, as a multiple of the largest batch (raised when the sample says so)
        # overlap the owner's level-2 scatter with the exchange of the later chunks (chunk-wise host barrier)
        self.overlap = not os.environ.get("ORION_XCHG_NO_OVERLAP")
        self.flags = HostFlags(dist, self.coll) if self.fused == 3 and self.overlap else None
        self.d_send = None
        self.recv, self.recv_cap, self.peer_ptrs = None, 0, None
        self.geom = None
        self.t = {}

    # ---- receive buffer shared with the peers (collective) ----
    def _ensure_recv(self, need_keys):
        torch, dist = self.torch, self.dist
        need = torch.tensor([need_keys], dtype=torch.int64)
        need = int(self.coll.all_reduce(need.cuda() if self.coll.native else need, dist.ReduceOp.MAX).item())
        if need <= self.recv_cap:
            return
        self._release_recv()
        cap = int(need * 1.15) + 1024
        self.recv = self.ok.PeerBuffer(cap * 8)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.recv.handle_bytes())
        self.peer_ptrs = [self.recv.ptr if r == self.rank else self.ok.PeerBuffer.open(handles[r])
                          for r in range(self.world)]
        self.recv_cap = cap

    def _release_recv(self):
        if self.recv is None:
            return
        self.torch.cuda.synchronize()
        self.coll.barrier()
        for r, p in enumerate(self.peer_ptrs):
            if r != self.rank:
                self.ok.PeerBuffer.close_peer(p)
        self.coll.barrier()
        self.recv.destroy()
        self.recv, self.recv_cap, self.peer_ptrs = None, 0, None

    def _agree(self, ok_flag):
        """MIN over the ranks of a success flag: every rank takes the same branch (also a barrier)"""
        t = self.torch.tensor([int(ok_flag)], dtype=self.torch.int32)
        return bool(int(self.coll.all_reduce(t.cuda() if self.coll.native else t, self.dist.ReduceOp.MIN).item()))

    def _max(self, v):
        t = self.torch.tensor([int(v)], dtype=self.torch.int64)
        return int(self.coll.all_reduce(t.cuda() if self.coll.native else t, self.dist.ReduceOp.MAX).item())

    def count_batch_device(self, d_bases, n_bases, d_off, n_reads):
        if self.fused == 3:
            return self._count_xchg(d_bases, n_bases, d_off, n_reads)
        if self.fused == 2:
            return self._count_sharded(d_bases, n_bases, d_off, n_reads)
        if self.fused == 1:
            return self._count_fused(d_bases, n_bases, d_off, n_reads)
        return self._count_unfused(d_bases, n_bases, d_off, n_reads)

    def _count_xchg(self, d_bases, n_bases, d_off, n_reads):
        """sample -> [reduce-scatter of the fine histogram + all-gather of the per-chunk level-1 histograms] -> chunked
        scatter, one copy-engine peer copy per (owner, chunk) under the next chunk's extraction -> [agreement
        all-reduce: also the barrier] -> fills from the sub-block headers, level 2 chunk by chunk, count."""
        torch, W = self.torch, self.world
        t0 = time.perf_counter()
        nmax = self._max(n_bases)
        if self.geom is None or self.geom.get("mode") != 3 or nmax > self.geom["nmax"]:
            sub_bits, l1_bits, n_chunks, cap = self.counter.xchg_geometry(nmax)
            self._ensure_recv(cap)
            self.counter.shard_set_buffers(self.peer_ptrs, self.recv_cap)
            i32 = dict(dtype=torch.int32, device="cuda")
            self.geom = {"mode": 3, "nmax": nmax, "sub_bits": sub_bits, "l1_bits": l1_bits, "n_chunks": n_chunks,
                         "hist_fine": torch.empty(W << sub_bits, **i32), "hist_mine": torch.empty(1 << sub_bits, **i32),
                         "hist_l1c": torch.empty(n_chunks * (W << l1_bits), **i32),
                         "l1c_all": torch.empty(W * n_chunks * (W << l1_bits), **i32)}
        g = self.geom
        self.counter.xchg_sample_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads,
                                        g["hist_fine"].data_ptr(), g["hist_l1c"].data_ptr())
        self.coll.reduce_scatter(g["hist_mine"], g["hist_fine"])
        self.coll.all_gather(g["l1c_all"], g["hist_l1c"])
        h_l1c_all = g["l1c_all"].cpu().numpy().view(np.uint32)       # the layout is planned on the host, identically on every rank
        # what every owner will receive is known now (sampled): a skewed input (35 % GC: 1.6 x the mean on one of 8
        # owners) needs more room than the default margin -- agree on a larger one and take the geometry again
        recv_est = h_l1c_all.reshape(W, g["n_chunks"], W, -1).sum(axis=(0, 1, 3)).astype(np.float64) * 16.0
        need = float(recv_est.max()) * 1.08 / max(1, nmax)
        if need > self.margin:
            self.margin = min(8.0, need * 1.1)
            self.counter.shard_set_margin(self.margin)
            self.geom = None
            return self._count_xchg(d_bases, n_bases, d_off, n_reads)
        t1 = time.perf_counter()
        ok_flag = 1
        try:
            if self.flags is None:
                self.counter.xchg_scatter_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads,
                                                 g["hist_mine"].data_ptr(), h_l1c_all)
            else:
                began = True
                try:
                    self.counter.xchg_scatter_begin(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads,
                                                    g["hist_mine"].data_ptr(), h_l1c_all)
                except self.ok.OrionError:
                    began = False                   # (still walks the barriers below: the others are waiting in them)
                    raise
                finally:
                    for ch in range(g["n_chunks"]):
                        if began:
                            self.counter.xchg_chunk_sent(ch)     # my copies of chunk ch have landed
                        self.flags.barrier()                     # ... and everybody else's
                        if began:
                            self.counter.xchg_chunk_recv(ch)     # level 2 of chunk ch, under the exchange of the later chunks
                self.counter.xchg_scatter_end()
        except self.ok.OrionError as e:
            ok_flag = 0
            if os.environ.get("ORION_VERBOSE"):
                print(f"[rank {self.rank}] chunked exchange failed, falling back: {e}", flush=True)
        t2 = time.perf_counter()
        all_ok = self._agree(ok_flag)       # every sender's copies have landed in my buffer
        t3 = time.perf_counter()
        if not all_ok:                      # a sampled region overflowed somewhere: every rank recounts through the exact route
            self.fallbacks += 1
            self.counter.abort_batch()
            return self._count_unfused(d_bases, n_bases, d_off, n_reads)
        count_ok = 1
        try:
            self.counter.xchg_count_device()
        except self.ok.OrionError as e:
            count_ok = 0
            if os.environ.get("ORION_VERBOSE"):
                print(f"[rank {self.rank}] sharded count failed, recounting: {e}", flush=True)
        if not self._agree(count_ok):       # e.g. a capacity hint that is too low only shows once the shared-memory tables overflow
            self.fallbacks += 1
            self.counter.abort_batch()
            self.counter.set_capacity_hint(0)      # every rank drops the hint: the geometry must stay collective
            self.hinted, self.geom = False, None
            return self._count_unfused(d_bases, n_bases, d_off, n_reads)
        self.counter.commit_batch()         # every rank counted its share: merge into the result of earlier batches (if any)
        t4 = time.perf_counter()
        st = self.counter.stats()
        self.t = {"route_ms": (t2 - t0) * 1e3, "route_count_ms": (t1 - t0) * 1e3, "route_scatter_ms": (t2 - t1) * 1e3,
                  "exchange_ms": (t3 - t2) * 1e3, "count_ms": (t4 - t3) * 1e3,
                  "sent_kmers": float("nan"), "sent_off_rank": float("nan"), "recv_kmers": int(st["n_windows"])}

    def _count_sharded(self, d_bases, n_bases, d_off, n_reads):
        """sample -> [reduce-scatter + all-gather of the histograms] -> scatter into the owners' level-1
        regions over NVLink -> [all-gather of the cursors: also the barrier] -> level 2 + count."""
        torch, dist, W = self.torch, self.dist, self.world
        t0 = time.perf_counter()
        nmax = self._max(n_bases)
        if self.geom is None or self.geom.get("mode") != 2 or nmax > self.geom["nmax"]:
            sub_bits, l1_bits, cap = self.counter.shard_geometry(nmax)
            self._ensure_recv(cap)
            self.counter.shard_set_buffers(self.peer_ptrs, self.recv_cap)
            i32 = dict(dtype=torch.int32, device="cuda")
            self.geom = {"mode": 2, "nmax": nmax, "sub_bits": sub_bits, "l1_bits": l1_bits,
                         "hist_fine": torch.empty(W << sub_bits, **i32), "hist_l1": torch.empty(W << l1_bits, **i32),
                         "hist_mine": torch.empty(1 << sub_bits, **i32), "l1_all": torch.empty(W * (W << l1_bits), **i32),
                         "cursors": torch.empty(W << l1_bits, **i32), "cur_all": torch.empty(W * (W << l1_bits), **i32),
                         "flag": torch.empty(1, **i32)}
        g = self.geom
        self.counter.shard_sample_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads,
                                         g["hist_fine"].data_ptr(), g["hist_l1"].data_ptr())
        self.coll.reduce_scatter(g["hist_mine"], g["hist_fine"])
        self.coll.all_gather(g["l1_all"], g["hist_l1"])
        torch.cuda.current_stream().synchronize()
        t1 = time.perf_counter()
        ok_flag = 1
        try:
            self.counter.shard_scatter_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads,
                                              g["hist_mine"].data_ptr(), g["l1_all"].data_ptr(), g["cursors"].data_ptr())
        except self.ok.OrionError as e:
            ok_flag = 0
            if os.environ.get("ORION_VERBOSE"):
                print(f"[rank {self.rank}] sharded scatter failed, falling back: {e}", flush=True)
        t2 = time.perf_counter()
        all_ok = self._agree(ok_flag)
        self.coll.all_gather(g["cur_all"], g["cursors"])      # every sender has finished writing into my buffer
        t3 = time.perf_counter()
        if not all_ok:          # a sampled region overflowed somewhere: every rank recounts through the exact route
            self.fallbacks += 1
            self.counter.abort_batch()
            return self._count_unfused(d_bases, n_bases, d_off, n_reads)
        count_ok = 1
        try:
            self.counter.shard_count_device(g["cur_all"].data_ptr())
        except self.ok.OrionError as e:
            count_ok = 0
            if os.environ.get("ORION_VERBOSE"):
                print(f"[rank {self.rank}] sharded count failed, recounting: {e}", flush=True)
        # ALWAYS agreed (hint or not): a rank that raised alone would leave the others waiting in the next collective
        if not self._agree(count_ok):
            self.fallbacks += 1
            self.counter.abort_batch()
            self.counter.set_capacity_hint(0)      # every rank drops the hint: the geometry must stay collective
            self.hinted, self.geom = False, None
            return self._count_unfused(d_bases, n_bases, d_off, n_reads)
        self.counter.commit_batch()         # every rank counted its share: merge into the result of earlier batches (if any)
        t4 = time.perf_counter()
        st = self.counter.stats()
        self.t = {"route_ms": (t2 - t0) * 1e3, "route_count_ms": (t1 - t0) * 1e3, "route_scatter_ms": (t2 - t1) * 1e3,
                  "exchange_ms": (t3 - t2) * 1e3, "count_ms": (t4 - t3) * 1e3,
                  "sent_kmers": float("nan"), "sent_off_rank": float("nan"), "recv_kmers": int(st["n_windows"])}

    def _count_unfused(self, d_bases, n_bases, d_off, n_reads):
        torch = self.torch
        if self.d_send is None or self.d_send.numel() < n_bases:
            self.d_send = torch.empty(n_bases, dtype=torch.int64, device=d_bases.device)
        t0 = time.perf_counter()
        counts = self.counter.route_batch_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads,
                                                 self.world, self.d_send.data_ptr())
        t1 = time.perf_counter()
        recv, _ = exchange(self.dist, torch, self.d_send, counts)
        torch.cuda.current_stream().synchronize()
        t2 = time.perf_counter()
        self.counter.add_kmers_device(recv.data_ptr(), recv.numel())
        t3 = time.perf_counter()
        self.t = {"route_ms": (t1 - t0) * 1e3, "exchange_ms": (t2 - t1) * 1e3, "count_ms": (t3 - t2) * 1e3,
                  "sent_kmers": int(counts.sum()), "sent_off_rank": int(counts.sum() - counts[self.rank]),
                  "recv_kmers": int(recv.numel())}
        del recv

    def _count_fused(self, d_bases, n_bases, d_off, n_reads):
        torch, dist = self.torch, self.dist
        t0 = time.perf_counter()
        counts = self.counter.route_count_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads, self.world)
        # world x world matrix M[src][dst]
        mine = torch.from_numpy(counts.astype(np.int64)).cuda()
        M = torch.empty(self.world * self.world, dtype=torch.int64, device="cuda")
        self.coll.all_gather(M, mine)
        M = M.cpu().numpy().reshape(self.world, self.world)
        recv_total = M.sum(axis=0)
        self._ensure_recv(int(recv_total[self.rank]))
        t1 = time.perf_counter()
        # my slice in rank d's buffer starts after the slices of the lower-ranked senders
        offs = np.concatenate([np.zeros((1, self.world), np.int64), np.cumsum(M, axis=0)[:-1]])[self.rank]
        dst = [self.peer_ptrs[d] + 8 * int(offs[d]) for d in range(self.world)]
        self.counter.route_scatter_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads, dst, counts)
        t2 = time.perf_counter()
        self.coll.barrier()                  # every sender has finished writing into my buffer
        t3 = time.perf_counter()
        self.counter.add_kmers_device(self.recv.ptr, int(recv_total[self.rank]))
        t4 = time.perf_counter()
        self.t = {"route_ms": (t1 - t0) * 1e3 + (t2 - t1) * 1e3, "route_count_ms": (t1 - t0) * 1e3,
                  "route_scatter_ms": (t2 - t1) * 1e3, "exchange_ms": (t3 - t2) * 1e3, "count_ms": (t4 - t3) * 1e3,
                  "sent_kmers": int(counts.sum()), "sent_off_rank": int(counts.sum() - counts[self.rank]),
                  "recv_kmers": int(recv_total[self.rank])}

    def clear(self):
        self.counter.clear()

    def close(self):
        self._release_recv()
        if self.flags is not None:
            self.flags.close()
            self.flags = None
        self.counter.close()


def all_vs_all(ok, torch, dist, sets, device=None):
    """All-vs-all intersection sizes over the ranks of a process group (BASELINE.json configs[4]).
    Every rank holds all the sets (256 x 5 M keys = 10 GB); the pairs are dealt round robin to the ranks and ONE
    all-reduce(sum) of the n x n matrix is the only exchange the path has.  `device`: where the matrix is reduced
    ("cuda" under NCCL, "cpu" under gloo)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes, upper = ok.all_vs_all_part(sets, rank, world)
    if device is None:
        device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    m = torch.from_numpy(upper.view(np.int64)).to(device)
    dist.all_reduce(m, op=dist.ReduceOp.SUM)
    return sizes, ok.finish_all_vs_all(sizes, m.cpu().numpy().view(np.uint64))


# ------------------------------------------------------------------ multi-GPU set algebra --
# build (e-2): genome i is built on rank i % world -- independent units, no exchange (build.rs:93-116).
# compare / query / classify (e-3, e-4): every set is cut at the SAME owner boundaries (key ranges of the canonical
# k-mer position, the rule of the sharded count) and slice r of every set travels to rank r in one all-to-all.
# A key lives in exactly one shard, so |A n B|, per-read hits and (matched, depth) are plain sums over the ranks:
# one all-reduce(sum) of integers completes each of them (compare.rs:51-66, query.rs:77-109, classify.rs:224-277).

def build_sets(ok, dist, k, genomes, make_batch):
    """genome-per-GPU build: -> {global index: KmerSet} of the genomes this rank owns.  make_batch(i) -> (bases, offsets)
    of genome i (host arrays, already normalised)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = {}
    own = list(range(rank, genomes, world))
    for c in range(0, len(own), 16):             # a few files at a time through ok_sets_build_many (builders side by side)
        chunk = own[c:c + 16]
        for i, s in zip(chunk, ok.KmerSet.build_many(k, [make_batch(i) for i in chunk])):
            mine[i] = s
    return mine


def reshard_sets(ok, torch, dist, k, mine, n_sets, key_device="cuda"):
    """mine: {global index: KmerSet} whole sets held by this rank (every index on exactly one rank).
    -> (shards, sizes): n_sets KmerSets holding THIS rank's key range of every set (global index order), and the
    full sizes of all sets (np.uint64[n_sets]).  One all-to-all of the keys is the only exchange.
    key_device: where the send / receive buffers of the keys live -- always "cuda" with the library; the CPU tests of
    this bookkeeping (gloo, a numpy stand-in for the sets) pass "cpu"."""
    rank, world = dist.get_rank(), dist.get_world_size()
    coll = Coll(dist, torch)
    dev = "cuda" if coll.native else "cpu"
    idx = sorted(mine)
    owners = [None] * world
    dist.all_gather_object(owners, idx)
    owner_of = {i: q for q, lst in enumerate(owners) for i in lst}
    assert sorted(owner_of) == list(range(n_sets)), "every set must live on exactly one rank"
    part = np.zeros((n_sets, world), dtype=np.int64)          # part[i][r] = keys of set i owned by rank r
    bounds = {}
    for i in idx:
        bnd = mine[i].shard_bounds(world).astype(np.int64)
        bounds[i] = bnd
        part[i] = bnd[1:] - bnd[:-1]
    t = torch.from_numpy(part).to(dev)
    coll.all_reduce(t)
    part = t.cpu().numpy()
    # send buffer: for every destination rank, its slice of each of my sets (ascending set index)
    send_counts = [int(sum(part[i][r] for i in idx)) for r in range(world)]
    send = torch.empty(max(1, sum(send_counts)), dtype=torch.int64, device=key_device)
    at = 0
    for r in range(world):
        for i in idx:
            n = int(part[i][r])
            if n:
                mine[i].copy_keys_device(int(bounds[i][r]), n, send.data_ptr() + 8 * at)
            at += n
    recv, recv_counts = exchange(dist, torch, send, send_counts)
    if key_device == "cuda":
        torch.cuda.synchronize()  # the library reads `recv` on its own streams: the collective must have landed first
    seg = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)      # where source rank q's keys start
    shards, used = [], [0] * world
    for i in range(n_sets):
        q, n = owner_of[i], int(part[i][rank])
        shards.append(ok.KmerSet.from_sorted_device(k, recv.data_ptr() + 8 * int(seg[q] + used[q]), n))
        used[q] += n
    assert all(used[q] == recv_counts[q] for q in range(world))
    return shards, part.sum(axis=1).astype(np.uint64)


def all_vs_all_sharded(ok, torch, dist, k, mine, n_sets, key_device="cuda"):
    """compare.rs:51-60 for every pair of n_sets sets spread over the ranks: identical key-range sharding of every
    set, every rank computes the whole n x n matrix over ITS key range, one all-reduce(sum) adds the ranges up.
    -> (sizes, full symmetric intersection matrix), on every rank."""
    coll = Coll(dist, torch)
    shards, sizes = reshard_sets(ok, torch, dist, k, mine, n_sets, key_device)
    _, upper = ok.all_vs_all_part(shards, 0, 1)
    m = torch.from_numpy(upper.view(np.int64).copy()).to("cuda" if coll.native else "cpu")
    coll.all_reduce(m)
    for s in shards:
        s.close()
    return sizes, ok.finish_all_vs_all(sizes, m.cpu().numpy().view(np.uint64))


def query_sharded(ok, torch, dist, shard_union, bases, off, norm_mode=1):
    """query.rs:77-109 with the database sharded by key range and the reads replicated: every rank counts the windows
    whose k-mer lies in ITS shard, one all-reduce(sum) of the per-read hits completes them."""
    coll = Coll(dist, torch)
    hits = shard_union.probe_reads(bases, off, norm_mode).astype(np.int64)
    t = torch.from_numpy(hits).to("cuda" if coll.native else "cpu")
    coll.all_reduce(t)
    return t.cpu().numpy().astype(np.uint32)


def classify_sharded(ok, torch, dist, shards, kmers, counts):
    """classify.rs:224-277 numerators and denominators with every reference sharded by key range: matched / depth per
    reference and |R| are sums over the ranks (the input count map is replicated).
    -> (matched[n_refs], depth[n_refs], ref_sizes[n_refs]) on every rank."""
    coll = Coll(dist, torch)
    m, d = ok.probe_counts_many(shards, kmers, counts)
    sizes = np.array([len(s) for s in shards], dtype=np.int64)
    t = torch.from_numpy(np.concatenate([m.astype(np.int64), d.astype(np.int64), sizes])).to("cuda" if coll.native else "cpu")
    coll.all_reduce(t)
    out = t.cpu().numpy().astype(np.uint64)
    n = len(shards)
    return out[:n], out[n:2 * n], out[2 * n:]


def verify_sharded_table(ok, dist, sc, K, bases, off, n_windows, slice_checker=None):
    """Parity of the sharded result, every rank: keys strictly ascending, rank boundaries in order (so the ranks'
    outputs concatenate to the sorted global table), sum of counts == windows received; and -- slice_checker given --
    one narrow key slice straddling the boundary between two ranks against an independent CPU pass over ALL ranks'
    reads (each rank checks its own reads, the partial tables are summed on rank 0).  -> dict on rank 0, None elsewhere.
    slice_checker(k, bases, offsets, lo, hi) -> (keys, counts) restricted to lo <= key <= hi: bench.py passes the
    oracle's; this package itself never loads the oracle."""
    rank, world = dist.get_rank(), dist.get_world_size()
    pk, pc, n = sc.counter.finish_raw(1)
    keys = np.ctypeslib.as_array(pk, shape=(max(n, 1),))[:n]
    counts = np.ctypeslib.as_array(pc, shape=(max(n, 1),))[:n]
    local = {"rank": rank, "n": int(n), "ascending": bool(n < 2 or np.all(keys[1:] > keys[:-1])),
             "first": int(keys[0]) if n else None, "last": int(keys[-1]) if n else None,
             "sum_counts": int(counts.sum(dtype=np.uint64)), "windows": int(n_windows)}
    infos = [None] * world
    dist.all_gather_object(infos, local)
    # slice: 1/4096 of the key space around the first key of the middle rank
    span = (1 << (2 * K)) >> 12
    mid = infos[world // 2]["first"] or 0
    lo, hi = max(0, mid - span // 2), mid + span // 2
    a, b = np.searchsorted(keys, np.uint64(lo), "left"), np.searchsorted(keys, np.uint64(hi), "right")
    part = {"gpu": (keys[a:b].copy(), counts[a:b].copy())}
    if slice_checker is not None:
        t0 = time.perf_counter()
        part["cpu"] = slice_checker(K, bases, off, lo, hi)
        part["cpu_s"] = time.perf_counter() - t0
    sc.counter.free_result(pk, pc)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(part, parts, dst=0)
    if rank != 0:
        return None
    out = {"ascending_all_ranks": all(i["ascending"] for i in infos),
           "rank_boundaries_in_order": all(infos[r]["last"] < infos[r + 1]["first"] for r in range(world - 1)
                                           if infos[r]["n"] and infos[r + 1]["n"]),
           "sum_counts_eq_windows": all(i["sum_counts"] == i["windows"] for i in infos),
           "distinct_total": sum(i["n"] for i in infos), "windows_total": sum(i["windows"] for i in infos)}
    if slice_checker is not None:
        gk = np.concatenate([p["gpu"][0] for p in parts])
        gc = np.concatenate([p["gpu"][1] for p in parts])
        ck = np.concatenate([p["cpu"][0] for p in parts])
        cc = np.concatenate([p["cpu"][1] for p in parts])
        uk, inv = np.unique(ck, return_inverse=True)
        uc = np.zeros(len(uk), dtype=np.uint64)
        np.add.at(uc, inv, cc)
        out["parity_slice_ok"] = bool(np.array_equal(gk, uk) and np.array_equal(gc, uc))
        out["slice"] = {"key_lo": int(lo), "key_hi": int(hi), "distinct": int(len(uk)), "windows": int(uc.sum()),
                        "straddles_ranks": [world // 2 - 1, world // 2] if world > 1 else [0],
                        "cpu_seconds_max": max(p["cpu_s"] for p in parts)}
    return out


def sub_batches(n_reads):
    """A rank's share may be more than one pass can take (32-bit offsets inside a batch): sub-batches of <= 11 M reads,
    each exchanged and counted on its own and merged into the rank's shard (reads are 150 bases: cuts at multiples
    of 8 reads keep the 16-byte alignment).  -> [(first read, end read), ...]"""
    MAXR = 11_000_000
    n_sub = (n_reads + MAXR - 1) // MAXR
    per = ((n_reads + n_sub - 1) // n_sub + 7) // 8 * 8
    return [(a, min(n_reads, a + per)) for a in range(0, n_reads, per)]


def bench_config(workload_config, config, n_reads, genome_len, world):
    """the `config` object of a sharded bench line (n_reads per rank, genome_len of the whole job): bench.py's reference
    arm prints the same one"""
    extra = {}
    if config == 3:
        extra = {"workload": f"count canonical 31-mers from {n_reads * world}x150bp synthetic reads key-range-sharded over "
                             f"{world} GPUs (BASELINE.json configs[2])", "total_reads": n_reads * world}
    return dict(workload_config(n_reads, genome_len, world), sub_batches_per_rank=len(sub_batches(n_reads)), **extra)


def bench(args, ok, synth, torch, world, rank, local, make_workload, workload_config, ClockSampler, measured_peak,
          metric, numa_node=None, slice_checker=None, numa_note=None):
    import torch.distributed as dist
    K = 31
    n_reads = args.reads
    genome_len = (args.genome or n_reads * 5) * world         # coverage stays 30x as ranks are added
    g, bases, off = make_workload(ok, synth, n_reads, genome_len, first_read=rank * n_reads)
    n_bases = len(bases)
    h_bases = torch.from_numpy(bases)          # page-locked (ok_host_alloc) host buffers
    h_off = torch.from_numpy(off.view(np.int64))
    d_bases = h_bases.cuda()
    d_off = h_off.cuda()
    hint = args.hint or int(n_bases * 0.17)       # expected distinct k-mers per rank (30x coverage, 0.5 % errors)
    sc = ShardedCounter(ok, torch, dist, K, fused=int(os.environ.get("ORION_FUSED", "3")), capacity_hint=hint)

    cuts = sub_batches(n_reads)

    def count_all():
        for a, b in cuts:
            sc.count_batch_device(d_bases[a * 150:b * 150], (b - a) * 150, d_off, b - a)

    def step_device():
        sc.clear()
        count_all()
        return sc.counter.finish_device(1)

    def step_host():
        sc.clear()
        d_bases.copy_(h_bases, non_blocking=True)
        d_off.copy_(h_off, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        count_all()
        pk, pc, n = sc.counter.finish_raw(1)
        sc.counter.free_result(pk, pc)
        return n

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        acc = []
        for _ in range(steps):
            fn()
            acc.append(dict(sc.t, **{k: v for k, v in sc.counter.stats().items() if k.startswith("ms_")}))
        torch.cuda.synchronize()
        dist.barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item()), acc

    sampler = ClockSampler(local)
    launches0 = ok.launch_count()
    sampler.start()
    dt, acc = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop()
    launches = ok.launch_count() - launches0
    st = sc.counter.stats()
    dt_e2e, _ = timed(step_host, args.steps, max(1, args.warmup))
    n_out = sc.counter.finish_device(1)[2]

    # ---- parity of what was just timed (same counter, same geometry, same hint) -------------------------
    # The CPU slice pass costs ~0.2 us per base and thread; when the whole batch would take too long on this
    # host's cores the step is repeated on a prefix of every rank's reads (same buffers and geometry) and THAT
    # table is checked -- the line says which.
    parity = None
    if not getattr(args, "no_parity", False):
        threads = max(1, (os.cpu_count() or 1) // world)
        budget_s = float(getattr(args, "parity_seconds", 60.0))
        est_s = n_bases * 0.2e-6 / threads
        v_reads = n_reads if est_s <= budget_s else max(10_000, int(n_reads * budget_s / est_s))
        v_bases = v_reads * 150
        if v_reads != n_reads:
            v_reads = min(v_reads, cuts[0][1])
            v_bases = v_reads * 150
            sc.clear()
            sc.count_batch_device(d_bases, v_bases, d_off, v_reads)
        vb, vo = bases[:v_bases], off[:v_reads + 1]
        checker = (lambda k, b, o, lo, hi: slice_checker(k, b, o, lo, hi, threads)) if slice_checker else None
        parity = verify_sharded_table(ok, dist, sc, K, vb, vo, sc.counter.stats()["n_windows"], checker)
        if rank == 0:
            parity["reads_per_rank_checked"] = int(v_reads)
            parity["full_batch"] = bool(v_reads == n_reads)
            parity["cpu_threads_per_rank"] = threads

    # whole-job totals
    tot = torch.tensor([st["n_windows"], st["n_distinct"], n_bases], device="cuda", dtype=torch.float64)
    dist.all_reduce(tot)
    windows, distinct, total_bases = (float(x) for x in tot.tolist())
    mean = {k: float(np.mean([a[k] for a in acc])) for k in acc[0]}
    if rank == 0:
        peak, peak_src = measured_peak()
        alg_step = total_bases * 1.5 + windows * 16.0 + distinct * 32.0 + windows * 32.0   # + multi-GPU 32 W (8d)
        off_rank = mean["sent_off_rank"]
        if not np.isfinite(off_rank):       # the sharded scatter does not count per owner: the prior balances the owners
            off_rank = windows / world * (world - 1) / world
            mean["sent_off_rank"] = off_rank
            mean["sent_kmers"] = windows / world
        nvlink_bytes = off_rank * 8.0
        line = {
            "metric": metric, "value": total_bases / dt, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong" if getattr(args, "config", 2) == 3 else "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": bench_config(workload_config, getattr(args, "config", 2), n_reads, genome_len, world),
            "e2e": {"value": total_bases / dt_e2e, "unit": "bases/s", "ms_per_step": dt_e2e * 1e3,
                    "h2d_bytes_per_step": int((n_bases + (n_reads + 1) * 8) * world),
                    "d2h_bytes_per_step": int(16 * distinct), "rank0_numa_node": numa_node, "rank0_numa_note": numa_note},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole step (route + all-to-all + sharded count), rank 0 phases below",
                         "achieved": alg_step / dt / 1e9, "peak": peak * world, "unit": "GB/s",
                         "frac": alg_step / dt / 1e9 / (peak * world), "peak_source": peak_src + " x n_gpus",
                         "traffic": None,
                         "nvlink": {"bytes_sent_per_rank": nvlink_bytes,
                                    "transfer_ms": mean.get("route_scatter_ms", mean["exchange_ms"]),
                                    "achieved_GBs_per_rank": nvlink_bytes / (mean.get("route_scatter_ms", mean["exchange_ms"]) / 1e3) / 1e9,
                                    "peak_GBs": 770.0, "peak_source": "B200_PROFILING.md peer copy per direction",
                                    "mode": {3: "chunked exchange: k_part_scatter_bases builds per-chunk sub-blocks, one copy-engine peer copy per (owner, chunk) under the next chunk's extraction",
                                             2: "sharded scatter: k_part_scatter_bases<PEER> writes level-1 partitioned k-mers into the owners' buffers",
                                             1: "two-pass route fused into k_part_scatter_bases<PEER> (writes into peer memory)",
                                             0: "NCCL all_to_all_single"}[sc.fused]}},
            "phases_ms": mean,
            "table": {"windows": int(windows), "distinct": int(distinct), "rank0_distinct": int(n_out),
                      "spilled": int(st["n_spilled"]), "rank0_fallbacks_to_nccl_all_to_all": int(getattr(sc, "fallbacks", 0))},
            "cpu_baseline": None,
            "parity": parity,
            "parity_slice_ok": None if parity is None else parity.get("parity_slice_ok"),
        }
        print(json.dumps(line))
    sc.close()
    dist.barrier()
    dist.destroy_process_group()
