"""bench_sets.py -- the set-operation configurations of BASELINE.json (configs[3] and configs[4]) for bench.py.

  configs[3]  build a k-mer database from 1,000 synthetic 5 Mbp genomes and query containment of a 1 M-read sample
  configs[4]  all-vs-all Jaccard compare of 256 synthetic genome k-mer sets (k = 21)

Workloads follow SURVEY.md 8(d): 20 seeded ancestors (seeds 100..119), 50 descendants each at 0.1 % .. 5 %
substitution divergence; the query sample is drawn with the configs[1] read recipe from 10 of the genomes plus 10 %
reads from an unrelated genome.  One JSON line per run, same keys as the count configurations.
"""
import json
import os
import time

import numpy as np

GENOME_LEN = 5_000_000
PER_ANCESTOR = 50


def genome(synth, i, length=GENOME_LEN, cache={}):
    """genome i of the configs[3]/[4] collection (deterministic)"""
    a = i // PER_ANCESTOR
    key = (a, length)
    if key not in cache:
        cache.clear()                               # one ancestor at a time is enough: genomes are generated in order
        cache[key] = synth.genome(100 + a, length)
    j = i % PER_ANCESTOR
    return synth.mutate(cache[key], 1000 + i, 1000 + j * 1000)      # 0.1 % .. 5 % substitutions


def query_sample(synth, n_reads, n_genomes, length=GENOME_LEN):
    """n_reads x 150 bp: 90 % from 10 genomes of the collection, 10 % from an unrelated genome"""
    picks = [int(x) for x in np.linspace(0, n_genomes - 1, 10)]
    per = int(n_reads * 0.9) // len(picks)
    parts = [synth.reads(genome(synth, g, length), 500 + t, per) for t, g in enumerate(picks)]
    rest = n_reads - per * len(picks)
    parts.append(synth.reads(synth.genome(999, length), 599, rest))
    return np.concatenate(parts), synth.read_offsets(n_reads)


def _clock(fn, steps, warmup, sync):
    for _ in range(warmup):
        fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    sync()
    return (time.perf_counter() - t0) / steps


# ------------------------------------------------------------------------------------ configs[4] --
def run_compare(args, ctx):
    """all-vs-all of n_sets genome k-mer sets, k = 21.  value: the sets resident in HBM -> sizes + the full intersection
    matrix (compare.rs:51-66 for every pair; Jaccard follows on the host).  e2e: sorted host arrays in (as loaded from
    a .db file), matrix out.  N > 1: every set cut at the same key-range boundaries, every rank computes the whole
    matrix over its range, one all-reduce(sum) (multi.all_vs_all_sharded)."""
    ok, synth, torch = ctx["ok"], ctx["synth"], ctx["torch"]
    world, rank = ctx["world"], ctx["rank"]
    k, n_sets = 21, args.sets
    length = args.genome_len
    t_gen = time.perf_counter()
    mine = {}
    for i in range(n_sets):
        if i % world != rank:
            continue
        g = genome(synth, i, length)
        s = ok.KmerSet.build(k)
        s.add_batch(g, np.array([0, len(g)], np.uint64))
        len(s)
        mine[i] = s
    t_gen = time.perf_counter() - t_gen
    launches0 = ok.launch_count()
    sampler = ctx["ClockSampler"](ctx["local"])
    if world == 1:
        sets = [mine[i] for i in range(n_sets)]
        out = {}

        def step():
            out["r"] = ok.all_vs_all(sets)
        sampler.start()
        dt = _clock(step, args.steps, args.warmup, torch.cuda.synchronize)
        clocks = sampler.stop()
        sizes, inter = out["r"]
        launches = (ok.launch_count() - launches0) // (args.steps + args.warmup)
        # e2e: the sets arrive as sorted host arrays (a loaded .db), the matrix goes back to the host
        host = [s.to_array() for s in sets]

        def step_host():
            tmp = [ok.KmerSet.from_sorted(k, a) for a in host]
            r = ok.all_vs_all(tmp)
            for t in tmp:
                t.close()
            return r
        dt_e2e = _clock(step_host, max(1, args.steps // 2), 1, torch.cuda.synchronize)
        h2d = int(sum(len(a) for a in host) * 8)
    else:
        import torch.distributed as dist
        from orion_kmer_b200 import multi
        out = {}

        def step():
            out["r"] = multi.all_vs_all_sharded(ok, torch, dist, k, mine, n_sets)

        def sync():
            torch.cuda.synchronize()
            dist.barrier()
        sampler.start()
        dt = _clock(step, args.steps, args.warmup, sync)
        clocks = sampler.stop()
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        sizes, inter = out["r"]
        launches = (ok.launch_count() - launches0) // (args.steps + args.warmup)
        dt_e2e, h2d = dt, 0          # the sets are built where they are compared: no host leg in the sharded job
    n_pairs = n_sets * (n_sets - 1) // 2
    total_keys = int(np.asarray(sizes, dtype=np.uint64).sum())
    # sorted-merge model of SURVEY 8(d): 8 (|A| + |B|) bytes per pair
    alg = 8.0 * float(sum(int(sizes[i]) * (n_sets - 1) for i in range(n_sets)))
    peak, peak_src = ctx["measured_peak"]()
    # Which form ran?  The keyed pass (setops.cuh) is a handful of launches per step, the row form two per row.
    # keyed: every key of every set is read once (+ the tile bounds); rows: SURVEY 8(d)'s sorted-merge model per pair.
    keyed = launches < 2 * (n_sets - 1) + (n_sets if world > 1 else 0)      # (N > 1: + one order check per received shard)
    one_pass = 8.0 * total_keys + 4.0 * n_sets * (total_keys / 3072.0 + 1.0)
    pairwise = {"bytes": alg, "achieved": alg / dt / 1e9, "frac": alg / dt / 1e9 / (peak * world),
                "note": "SURVEY 8(d)'s sorted-merge model, 8(|A|+|B|) per pair: what a pair-by-pair implementation would move"}
    if keyed:
        roofline = {"bound": "hbm", "kernel": "k_ava_tiles (keyed all-vs-all, setops.cuh: every key of every set read once; per tile a hashed "
                                              "shared-memory table, bit rows of the shared keys, AND + POPC per 8 x 8 pair block)",
                    "achieved": one_pass / dt / 1e9, "peak": peak * world, "unit": "GB/s", "frac": one_pass / dt / 1e9 / (peak * world),
                    "peak_source": peak_src + (" x n_gpus" if world > 1 else ""), "traffic": None, "algorithmic_bytes_per_launch": one_pass,
                    "pairwise_model": pairwise,
                    "note": "algorithmic bytes of the keyed pass: 8 B per key + 4 B per (set, tile) bound; the kernel is bound by shared-memory "
                            "atomics and issue (AND / POPC), not by HBM. pairwise_model keeps round 1's figure so that rounds compare "
                            "(its frac can exceed 1: the keyed pass does not read a set once per pair)"}
    else:
        roofline = {"bound": "hbm", "kernel": "k_intersect_row_tiled (row of the all-vs-all: A tile in registers, B ranges through shared memory)",
                    "achieved": pairwise["achieved"], "peak": peak * world, "unit": "GB/s", "frac": pairwise["frac"],
                    "peak_source": peak_src + (" x n_gpus" if world > 1 else ""), "traffic": None, "algorithmic_bytes_per_launch": alg,
                    "note": "sorted-merge model 8(|A|+|B|) per pair; the kernel is shared-memory-search bound, not HBM bound"}
    line = None
    if rank == 0:
        # parity: a sample of pairs against the oracle's compare (compare.rs:51-66), Jaccard from the integers
        chk = ctx["oracle_compare_sample"](k, n_sets, length, sizes, inter, args.parity_pairs)
        jac = [inter[i, j] / max(1, int(sizes[i]) + int(sizes[j]) - int(inter[i, j])) for i, j in ((0, 1), (0, n_sets - 1))]
        line = {
            "metric": "set pairs compared per second (k=21 all-vs-all: |A|, |B|, |A n B| per pair; Jaccard on the host)",
            "value": n_pairs / dt, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"all-vs-all Jaccard compare of {n_sets} synthetic genome k-mer sets, k=21 (BASELINE.json configs[4])",
                       "k": k, "sets": n_sets, "genome_len": length, "pairs": n_pairs, "keys_total": total_keys,
                       "sharding": "key-range shards of every set, full matrix per shard, one all-reduce" if world > 1 else "none",
                       "l2": "40 MB sets against a 126 MB L2: the row's set stays resident, the others stream from HBM"},
            "e2e": {"value": n_pairs / dt_e2e, "unit": "pairs/s", "ms_per_step": dt_e2e * 1e3, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(n_sets * n_sets * 8)},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roofline,
            "example_jaccard": jac, "build_seconds_untimed": t_gen,
            "cpu_baseline": chk["cpu_baseline"], "parity_pairs_ok": chk["ok"], "parity_pairs_checked": chk["pairs"],
        }
    for s in mine.values():
        s.close()
    return line


# ------------------------------------------------------------------------------------ configs[3] --
def run_build_query(args, ctx):
    """build the sets of n_genomes genomes (build.rs:93-116), unify them (db_types.rs:43-48) and count, per read of the
    sample, the windows whose canonical k-mer is in the union (query.rs:77-109).  N > 1: genome-per-GPU build, the
    sets resharded by key range, reads replicated, all-reduce of the per-read hits."""
    ok, synth, torch = ctx["ok"], ctx["synth"], ctx["torch"]
    world, rank = ctx["world"], ctx["rank"]
    k, n_genomes, length, n_reads = 31, args.genomes, args.genome_len, args.query_reads
    own = [i for i in range(n_genomes) if i % world == rank]
    # inputs resident in HBM: this rank's genomes back to back (16-byte aligned: the genome length is a multiple of 16)
    assert length % 16 == 0
    h_all = ctx["pinned"](len(own) * length)
    for t, i in enumerate(own):
        h_all[t * length:(t + 1) * length] = genome(synth, i, length)
    d_all = torch.from_numpy(h_all).cuda()
    d_off = torch.tensor([0, length], dtype=torch.int64, device="cuda")
    reads, roff = query_sample(synth, n_reads, n_genomes, length)
    h_reads = ctx["pinned"](len(reads)); h_reads[:] = reads
    d_reads = torch.from_numpy(h_reads).cuda()
    d_roff = torch.from_numpy(roff.view(np.int64)).cuda()
    d_hits = torch.zeros(n_reads, dtype=torch.int32, device="cuda")
    off1 = np.array([0, length], np.uint64)
    state = {}
    if world > 1:
        import torch.distributed as dist
        from orion_kmer_b200 import multi

    def build(device):
        if not os.environ.get("ORION_BENCH_BUILD_SEQ"):      # (A/B knob: one file at a time, as round 1 did)
            # the files are independent units: ok_sets_build_many(_device) drives several builders side by side
            if device:
                lst = ok.KmerSet.build_many_device(k, [d_all.data_ptr() + t * length for t in range(len(own))], [length] * len(own),
                                                   [d_off.data_ptr()] * len(own), [1] * len(own))
            else:
                lst = ok.KmerSet.build_many(k, [(h_all[t * length:(t + 1) * length], off1) for t in range(len(own))])
            return dict(zip(own, lst))
        sets = {}
        for t, i in enumerate(own):
            s = ok.KmerSet.build(k)
            if device:
                s.add_batch_device(d_all.data_ptr() + t * length, length, d_off.data_ptr(), 1)
            else:
                s.add_batch(h_all[t * length:(t + 1) * length], off1)
            len(s)
            sets[i] = s
        return sets

    def step(device=True):
        for s in state.get("sets", {}).values():
            s.close()
        for s in state.get("extra", []):
            s.close()
        t0 = time.perf_counter()
        sets = build(device)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if world == 1:
            union = ok.KmerSet.union([sets[i] for i in range(n_genomes)])
            n_union = len(union)
            extra = [union]
        else:
            shards, _ = multi.reshard_sets(ok, torch, dist, k, sets, n_genomes)
            union = ok.KmerSet.union(shards)
            n_union = len(union)
            extra = shards + [union]
        t2 = time.perf_counter()
        if device:
            union.probe_reads_device(d_reads.data_ptr(), len(reads), d_roff.data_ptr(), n_reads, d_hits.data_ptr(), ok.RAW)
            hits = d_hits
        else:
            hits = torch.from_numpy(union.probe_reads(h_reads, roff, ok.RAW).astype(np.int32)).cuda()
        if world > 1:
            dist.all_reduce(hits)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        state.update(sets=sets, extra=extra, hits=hits.cpu().numpy().astype(np.uint32), n_union=n_union,
                     t=(t1 - t0, t2 - t1, t3 - t2))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    sampler = ctx["ClockSampler"](ctx["local"])
    launches0 = ok.launch_count()
    sampler.start()
    dt = _clock(step, args.steps, args.warmup, sync)
    clocks = sampler.stop()
    launches = (ok.launch_count() - launches0) // (args.steps + args.warmup)
    t_build, t_union, t_probe = state["t"]
    dt_e2e = _clock(lambda: step(False), max(1, args.steps // 2), 1, sync)
    if world > 1:
        t = torch.tensor([dt, dt_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e = (float(x) for x in t.tolist())
        tot = torch.tensor([state["n_union"]], device="cuda", dtype=torch.int64)
        dist.all_reduce(tot)
        n_union = int(tot.item())
    else:
        n_union = state["n_union"]
    total_bases = n_genomes * length + len(reads)
    peak, peak_src = ctx["measured_peak"]()
    line = None
    if rank == 0:
        chk = ctx["oracle_query_sample"](k, n_genomes, length, reads, roff, state["hits"], args.parity_reads, args.cpu_genomes)
        windows_q = n_reads * (150 - k + 1)
        alg_build = n_genomes * length * (1.5 + 16.0) + 32.0 * n_genomes * length     # count model of 8(d) per genome, D ~ W ~ B
        alg_probe = len(reads) * 1.5 + windows_q * 8.0
        line = {
            "metric": "bases/sec through build + union + query (k=31)", "value": total_bases / dt, "unit": "bases/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"build k-mer database from {n_genomes} synthetic {length / 1e6:g} Mbp genomes and query containment of a "
                                   f"{n_reads}-read sample (BASELINE.json configs[3])",
                       "k": k, "genomes": n_genomes, "genome_len": length, "query_reads": n_reads, "union_kmers": n_union,
                       "sharding": "genome-per-GPU build, key-range shards for the query, all-reduce of the hits" if world > 1 else "none",
                       "l2": "5 GB of genomes, a multi-GB union: nothing fits the 126 MB L2"},
            "e2e": {"value": total_bases / dt_e2e, "unit": "bases/s", "ms_per_step": dt_e2e * 1e3,
                    "h2d_bytes_per_step": int(len(own) * length * world + len(reads)), "d2h_bytes_per_step": int(n_reads * 4)},
            "gpu_launches": int(launches), "clocks": clocks,
            "phases_ms": {"build": t_build * 1e3, "union": t_union * 1e3, "probe": t_probe * 1e3},
            "roofline": {"bound": "hbm", "kernel": "set build (partitioned count per genome): the dominant phase",
                         "achieved": alg_build / max(t_build, 1e-9) / 1e9 / world, "peak": peak, "unit": "GB/s",
                         "frac": alg_build / max(t_build, 1e-9) / 1e9 / world / peak, "peak_source": peak_src, "traffic": None,
                         "algorithmic_bytes_per_launch": alg_build / n_genomes,
                         "probe": {"algorithmic_bytes": alg_probe, "achieved": alg_probe / max(t_probe, 1e-9) / 1e9,
                                   "frac": alg_probe / max(t_probe, 1e-9) / 1e9 / peak,
                                   "note": "probe model of 8(d): B*1.5 + W*8; by merge: the batch's distinct k-mers against the sorted union, "
                                           "then per-read probes into a table of the matches (sets >= 2^26 keys)"}},
            "cpu_baseline": chk["cpu_baseline"], "parity_hits_ok": chk["ok"], "parity_reads_checked": chk["reads"],
        }
    for s in state.get("sets", {}).values():
        s.close()
    for s in state.get("extra", []):
        s.close()
    return line


def emit(line):
    if line is not None:
        print(json.dumps(line))
