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


def exchange(dist, torch, send, send_counts, group=None):
    """send: 1-D int64 tensor holding the k-mers for rank 0, then rank 1, ... (send_counts each).
    -> (recv tensor, recv_counts list)."""
    world = dist.get_world_size(group)
    sc = torch.as_tensor(np.asarray(send_counts, dtype=np.int64), device=send.device)
    rc = torch.empty(world, dtype=torch.int64, device=send.device)
    dist.all_to_all_single(rc, sc, group=group)
    recv_counts = [int(x) for x in rc.cpu().tolist()]
    recv = torch.empty(sum(recv_counts), dtype=torch.int64, device=send.device)
    dist.all_to_all_single(recv, send[:int(sum(send_counts))], output_split_sizes=recv_counts,
                           input_split_sizes=[int(x) for x in send_counts], group=group)
    return recv, recv_counts


class ShardedCounter:
    """KmerCounter sharded over the ranks of a process group.

    fused=True (default): the routing kernel writes each owner's k-mers straight into that owner's
    receive buffer (CUDA-IPC peer memory over NVLink); the only collectives left are a world x world
    count matrix and a barrier.  fused=False: bucket locally, NCCL all_to_all_single, then count."""

    def __init__(self, ok, torch, dist, k, norm_mode=0, fused=2, capacity_hint=0):
        self.ok, self.torch, self.dist = ok, torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        # capacity_hint = expected distinct k-mers of THIS rank's shard (0: none): sizes the sub-partitions for their
        # distinct keys (fewer bins per scatter level); a hint that proves too low costs one recount, then it is ignored
        self.counter = ok.KmerCounter(k, norm_mode, capacity_hint)
        self.hinted = capacity_hint > 0
        self.counter.set_shard(self.rank, self.world)
        # 2: sharded scatter (sample, then ONE extraction pass that writes level-1 partitioned k-mers into
        #    the owners' buffers); 1: two-pass fused route (count, then scatter by owner); 0: NCCL all-to-all
        self.fused = int(fused)
        self.d_send = None
        self.recv, self.recv_cap, self.peer_ptrs = None, 0, None
        self.geom = None
        self.t = {}

    # ---- receive buffer shared with the peers (collective) ----
    def _ensure_recv(self, need_keys):
        torch, dist = self.torch, self.dist
        need = torch.tensor([need_keys], device="cuda", dtype=torch.int64)
        dist.all_reduce(need, op=dist.ReduceOp.MAX)
        need = int(need.item())
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
        self.dist.barrier()
        for r, p in enumerate(self.peer_ptrs):
            if r != self.rank:
                self.ok.PeerBuffer.close_peer(p)
        self.dist.barrier()
        self.recv.destroy()
        self.recv, self.recv_cap, self.peer_ptrs = None, 0, None

    def count_batch_device(self, d_bases, n_bases, d_off, n_reads):
        if self.fused == 2:
            return self._count_sharded(d_bases, n_bases, d_off, n_reads)
        if self.fused == 1:
            return self._count_fused(d_bases, n_bases, d_off, n_reads)
        return self._count_unfused(d_bases, n_bases, d_off, n_reads)

    def _count_sharded(self, d_bases, n_bases, d_off, n_reads):
        """sample -> [reduce-scatter + all-gather of the histograms] -> scatter into the owners' level-1
        regions over NVLink -> [all-gather of the cursors: also the barrier] -> level 2 + count."""
        torch, dist, W = self.torch, self.dist, self.world
        t0 = time.perf_counter()
        nmax = torch.tensor([n_bases], device="cuda", dtype=torch.int64)
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX)
        nmax = int(nmax.item())
        if self.geom is None or nmax > self.geom["nmax"]:
            sub_bits, l1_bits, cap = self.counter.shard_geometry(nmax)
            self._ensure_recv(cap)
            self.counter.shard_set_buffers(self.peer_ptrs, self.recv_cap)
            i32 = dict(dtype=torch.int32, device="cuda")
            self.geom = {"nmax": nmax, "sub_bits": sub_bits, "l1_bits": l1_bits,
                         "hist_fine": torch.empty(W << sub_bits, **i32), "hist_l1": torch.empty(W << l1_bits, **i32),
                         "hist_mine": torch.empty(1 << sub_bits, **i32), "l1_all": torch.empty(W * (W << l1_bits), **i32),
                         "cursors": torch.empty(W << l1_bits, **i32), "cur_all": torch.empty(W * (W << l1_bits), **i32),
                         "flag": torch.empty(1, **i32)}
        g = self.geom
        self.counter.shard_sample_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads,
                                         g["hist_fine"].data_ptr(), g["hist_l1"].data_ptr())
        dist.reduce_scatter_tensor(g["hist_mine"], g["hist_fine"])
        dist.all_gather_into_tensor(g["l1_all"], g["hist_l1"])
        torch.cuda.current_stream().synchronize()
        t1 = time.perf_counter()
        ok_flag = 1
        try:
            self.counter.shard_scatter_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads,
                                              g["hist_mine"].data_ptr(), g["l1_all"].data_ptr(), g["cursors"].data_ptr())
        except self.ok.OrionError as e:
            ok_flag = 0
            self.fallbacks = getattr(self, "fallbacks", 0) + 1
            if os.environ.get("ORION_VERBOSE"):
                print(f"[rank {self.rank}] sharded scatter failed, falling back: {e}", flush=True)
        t2 = time.perf_counter()
        g["flag"].fill_(ok_flag)
        dist.all_reduce(g["flag"], op=dist.ReduceOp.MIN)
        dist.all_gather_into_tensor(g["cur_all"], g["cursors"])      # every sender has finished writing into my buffer
        all_ok = int(g["flag"].item())
        t3 = time.perf_counter()
        if not all_ok:          # a sampled region overflowed somewhere: every rank recounts through the exact route
            self.counter.clear()
            return self._count_unfused(d_bases, n_bases, d_off, n_reads)
        count_ok = 1
        try:
            self.counter.shard_count_device(g["cur_all"].data_ptr())
        except self.ok.OrionError:
            if not self.hinted:
                raise
            count_ok = 0
        if self.hinted:         # a capacity hint that is too low only shows once the shared-memory tables overflow
            g["flag"].fill_(count_ok)
            dist.all_reduce(g["flag"], op=dist.ReduceOp.MIN)
            if not int(g["flag"].item()):
                self.fallbacks = getattr(self, "fallbacks", 0) + 1
                self.counter.clear()
                self.counter.set_capacity_hint(0)      # every rank drops the hint: the geometry must stay collective
                self.hinted, self.geom = False, None
                return self._count_unfused(d_bases, n_bases, d_off, n_reads)
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
        dist.all_gather_into_tensor(M, mine)
        M = M.cpu().numpy().reshape(self.world, self.world)
        recv_total = M.sum(axis=0)
        self._ensure_recv(int(recv_total[self.rank]))
        t1 = time.perf_counter()
        # my slice in rank d's buffer starts after the slices of the lower-ranked senders
        offs = np.concatenate([np.zeros((1, self.world), np.int64), np.cumsum(M, axis=0)[:-1]])[self.rank]
        dst = [self.peer_ptrs[d] + 8 * int(offs[d]) for d in range(self.world)]
        self.counter.route_scatter_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads, dst, counts)
        t2 = time.perf_counter()
        dist.barrier()                       # every sender has finished writing into my buffer
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


def bench(args, ok, synth, torch, world, rank, local, make_workload, workload_config, ClockSampler, measured_peak,
          metric, numa_node=None):
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
    sc = ShardedCounter(ok, torch, dist, K, fused=int(os.environ.get("ORION_FUSED", "2")), capacity_hint=hint)

    def step_device():
        sc.clear()
        sc.count_batch_device(d_bases, n_bases, d_off, n_reads)
        return sc.counter.finish_device(1)

    def step_host():
        sc.clear()
        d_bases.copy_(h_bases, non_blocking=True)
        d_off.copy_(h_off, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        sc.count_batch_device(d_bases, n_bases, d_off, n_reads)
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
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(n_reads, genome_len, world),
            "e2e": {"value": total_bases / dt_e2e, "unit": "bases/s", "ms_per_step": dt_e2e * 1e3,
                    "h2d_bytes_per_step": int((n_bases + (n_reads + 1) * 8) * world),
                    "d2h_bytes_per_step": int(16 * distinct), "rank0_numa_node": numa_node},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole step (route + all-to-all + sharded count), rank 0 phases below",
                         "achieved": alg_step / dt / 1e9, "peak": peak * world, "unit": "GB/s",
                         "frac": alg_step / dt / 1e9 / (peak * world), "peak_source": peak_src + " x n_gpus",
                         "traffic": None,
                         "nvlink": {"bytes_sent_per_rank": nvlink_bytes,
                                    "transfer_ms": mean.get("route_scatter_ms", mean["exchange_ms"]),
                                    "achieved_GBs_per_rank": nvlink_bytes / (mean.get("route_scatter_ms", mean["exchange_ms"]) / 1e3) / 1e9,
                                    "peak_GBs": 770.0, "peak_source": "B200_PROFILING.md peer copy per direction",
                                    "mode": {2: "sharded scatter: k_part_scatter_bases<PEER> writes level-1 partitioned k-mers into the owners' buffers",
                                             1: "two-pass route fused into k_part_scatter_bases<PEER> (writes into peer memory)",
                                             0: "NCCL all_to_all_single"}[sc.fused]}},
            "phases_ms": mean,
            "table": {"windows": int(windows), "distinct": int(distinct), "rank0_distinct": int(n_out),
                      "spilled": int(st["n_spilled"]), "rank0_fallbacks_to_nccl_all_to_all": int(getattr(sc, "fallbacks", 0))},
            "cpu_baseline": None,
        }
        print(json.dumps(line))
    sc.close()
    dist.barrier()
    dist.destroy_process_group()
