#!/usr/bin/env python
"""bench.py -- bases/sec counted (canonical k=31), BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R] [--config 1..5]

  --config 2 (default)   BASELINE.json configs[1]: 10 M x 150 bp reads per GPU, k = 31 (N > 1: weak scaling, sharded)
  --config 1             configs[0]: one 5 Mbp genome, k = 21 (the reference's own CPU-runnable case)
  --config 3             configs[2]: --total-reads (default 100 M) x 150 bp split over the N GPUs (strong), N >= 2
  --config 4             configs[3]: build a database from --genomes genomes, query --query-reads reads (bench_sets.py)
  --config 5             configs[4]: all-vs-all compare of --sets genome k-mer sets, k = 21 (bench_sets.py)

A step = one whole count job over one batch of synthetic reads (BASELINE.json configs[1]:
10M x 150 bp reads, k = 31): empty table -> extract + count every window -> sorted
(k-mer, count) arrays.  `value` times it with the batch already resident in HBM and the result
left in HBM; `e2e` times the same job through the C ABI with HOST buffers (page-locked input,
host result arrays), copies inside the timed region.  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 31
READ_LEN = 150
GENOME_SEED, READ_SEED = 3, 3
METRIC = "bases/sec counted (canonical k=31)"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML polled every 5 ms from a
    thread (a 40 ms step is far shorter than nvidia-smi's start-up); nvidia-smi -lms as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.proc, self.lines = gpu, None, []
        self.nv, self.h, self.stop_flag, self.sm, self.reasons, self.mx = None, None, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nv:
            self.stop_flag = False
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.nv:
            self.stop_flag = True
            self.t.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml, 5 ms poll"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


NUMA_NOTE = {"why": "not attempted"}


def bind_to_gpu_numa_node(gpu):
    """Run this rank (and first-touch its page-locked buffers) on the CPUs next to its GPU: with 8 ranks
    on a two-socket host, buffers on the wrong node cross the socket link on every H2D / D2H copy.
    -> the node, or None (NUMA_NOTE["why"] then says what was found)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dev = "/sys/bus/pci/devices/" + bus.lower()[-12:]
        node = int(open(dev + "/numa_node").read())
        if node < 0:
            NUMA_NOTE["why"] = f"{dev}/numa_node reads {node}: the host exposes no NUMA node for this GPU"
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        NUMA_NOTE["why"] = f"bound to node {node} ({len(cpus)} CPUs)" if cpus else f"node {node} has no CPU this process may run on"
        return node
    except Exception as e:
        NUMA_NOTE["why"] = f"{type(e).__name__}: {e}"
        return None


def pinned_array(ok, nbytes, dtype):
    p = C.c_void_p()
    ok._check(ok.lib().ok_host_alloc(C.byref(p), nbytes))
    arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(dtype)
    return arr, p


def make_workload(ok, synth, n_reads, genome_len, first_read=0, pinned=True):
    g = synth.genome(GENOME_SEED, genome_len)
    if pinned:
        bases, pb = pinned_array(ok, n_reads * READ_LEN, np.uint8)
        off, po = pinned_array(ok, (n_reads + 1) * 8, np.uint64)
    else:
        bases, off = np.empty(n_reads * READ_LEN, np.uint8), np.empty(n_reads + 1, np.uint64)
    synth.reads(g, READ_SEED, n_reads, first_read=first_read, out=bases)
    synth.read_offsets(n_reads, out=off)
    return g, bases, off


# ------------------------------------------------------------------------------ reference arm --
def cpu_baseline(n_reads_total, genome_len, sample_reads, all_cores=True):
    """The CPU restatement of the reference (oracle/, kind 'port': the Rust crate cannot be built
    here) on a bounded sample of the same workload: faithful single thread (count.rs:68-79 is a
    sequential loop), plus the all-cores variant labelled as not reference behaviour."""
    import oracle
    import orion_kmer_b200 as ok
    from orion_kmer_b200 import synth
    oracle.build()
    sample_reads = min(sample_reads, n_reads_total)
    _, bases, off = make_workload(ok, synth, sample_reads, genome_len, pinned=False)
    t0 = time.perf_counter()
    keys, counts = oracle.count_batch(K, bases, off, 1, True)
    dt = time.perf_counter() - t0
    out = {"value": len(bases) / dt, "unit": "bases/s", "cores": 1, "kind": "port",
           "sample": f"first {sample_reads} reads ({len(bases)} bases) of the workload, {dt:.2f} s, "
                     f"{len(keys)} distinct",
           "host_cores_available": os.cpu_count(),
           "note": "a bounded sample, not BASELINE.md's full run: its table is several times smaller than the full job's and "
                   "friendlier to the CPU caches, so the CPU figure is if anything flattering (the speed-up is conservative)"}
    if all_cores:
        nt = min(os.cpu_count() or 1, 64)
        t0 = time.perf_counter()
        k2, c2 = oracle.count_batch_mt(K, bases, off, nt)
        dt2 = time.perf_counter() - t0
        assert np.array_equal(keys, k2) and np.array_equal(counts, c2)
        out["all_cores"] = {"value": len(bases) / dt2, "cores": nt,
                            "note": "NOT reference behaviour (its count loop is single-threaded)"}
    return out, (bases, off, keys, counts)


def oracle_slice_checker():
    """The checker multi.bench verifies one key slice of the sharded table with (the oracle is test infrastructure:
    it is only ever the checker, never the thing measured)."""
    import oracle
    oracle.build()
    return oracle.count_batch_slice_mt


def run_reference(args):
    """--impl reference, configs 2 and 3: the oracle port (the Rust crate cannot be built here) on a bounded sample of
    OUR arm's workload -- same genome, same reads, same `config` object -- with the threads the reference itself uses
    for this path: one (count.rs:68-79 is a sequential loop over the records)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    n_reads = args.reads
    genome_len = (args.genome or n_reads * 5) * (world if world > 1 else 1)      # as multi.bench: coverage stays 30x
    sample = args.sample_reads
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base, _ = cpu_baseline(n_reads, genome_len, sample, all_cores=False)
        if i >= args.warmup:
            vals.append(base["value"])
    v = float(np.mean(vals))
    sample_bases = min(sample, n_reads) * READ_LEN
    base["value"] = v
    if world > 1:
        from orion_kmer_b200 import multi
        config = multi.bench_config(workload_config, args.config, n_reads, genome_len, world)
    else:
        config = workload_config(n_reads, genome_len, 1)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "bases/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample_bases / v,
            "higher_is_better": True, "scaling": "strong" if args.config == 3 else "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": config,
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


CONFIG1 = {"workload": "count canonical 21-mers in a single synthetic 5 Mbp genome record (BASELINE.json configs[0])",
           "k": 21, "n_runs": 10, "lower_case": "1 %", "seed": 1, "min_count": 1,
           "l2": "80 MB of keys against a 126 MB L2: partly L2-resident; the step is launch-latency bound (~15 launches)"}


def run_reference_config1(args):
    """--impl reference --config 1: the oracle port over the whole 5 Mbp genome record (k = 21), one thread"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle
    from orion_kmer_b200 import synth
    oracle.build()
    g = synth.config1_genome()
    off = np.array([0, len(g)], np.uint64)
    ts = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        keys, _ = oracle.count_batch(21, g, off)
        if i >= args.warmup:
            ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    v = len(g) / dt
    base = {"value": v, "unit": "bases/s", "cores": 1, "kind": "port",
            "sample": f"the whole genome ({len(g)} bases), {dt:.2f} s, {len(keys)} distinct", "host_cores_available": os.cpu_count()}
    print(json.dumps({"impl": "reference", "metric": "bases/sec counted (canonical k=21)", "value": v, "unit": "bases/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": dict(CONFIG1, bases=int(len(g))),
                      "cpu_baseline": base, "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(n_reads, genome_len, n_gpus):
    return {"workload": f"count canonical 31-mers from {n_reads}x150bp synthetic Illumina-like reads "
                        f"per GPU (BASELINE.json configs[1]{' / configs[2] sharded' if n_gpus > 1 else ''})",
            "k": K, "reads_per_gpu": n_reads, "read_len": READ_LEN, "genome_len": genome_len,
            "substitution_rate": 0.005, "n_rate": 0.001, "seeds": [GENOME_SEED, READ_SEED], "min_count": 1,
            "l2": "batch and table are both far larger than the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------ our arm --
def run_ours(args):
    import torch
    import orion_kmer_b200 as ok
    from orion_kmer_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local)
    ok.init(local)

    if args.config == 1:
        return run_config1(args, ok, synth, torch, local)
    if args.config in (4, 5):
        return run_sets_config(args, torch, ok, synth, world, rank, local)
    if args.config == 3 and world == 1:
        raise SystemExit("--config 3 is the sharded job of BASELINE.json configs[2]: run it under torchrun with --gpus 2, 4 or 8")
    n_reads = args.reads
    genome_len = args.genome or n_reads * 5
    if world > 1:
        from orion_kmer_b200 import multi
        return multi.bench(args, ok, synth, torch, world, rank, local, make_workload, workload_config,
                           ClockSampler, measured_peak, METRIC, numa_node, slice_checker=oracle_slice_checker(),
                           numa_note=NUMA_NOTE["why"])

    g, bases, off = make_workload(ok, synth, n_reads, genome_len)
    n_bases = len(bases)
    d_bases = torch.from_numpy(bases).cuda()
    d_off = torch.from_numpy(off.view(np.int64)).cuda()
    hint = args.hint or int(n_bases * 0.17)       # ~error k-mers + genome; the table grows if short
    counter = ok.KmerCounter(K, ok.NORMALIZED, hint)

    def step_device():
        counter.clear()
        counter.add_batch_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads)
        return counter.finish_device(1)

    e2e_t = {"add_batch_ms": [], "finish_ms": []}

    def step_host():
        counter.clear()
        t_a = time.perf_counter()
        counter.add_batch_ptr(bases.ctypes.data, off.ctypes.data, n_reads)
        t_b = time.perf_counter()
        pk, pc, n = counter.finish_raw(1)
        t_c = time.perf_counter()
        counter.free_result(pk, pc)
        e2e_t["add_batch_ms"].append((t_b - t_a) * 1e3); e2e_t["finish_ms"].append((t_c - t_b) * 1e3)
        return n

    # ---- value: batch resident in HBM ---------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ok.launch_count()
    PH = ["ms_fill", "ms_insert", "ms_readout", "ms_sample", "ms_scatter1", "ms_scatter2", "ms_count", "ms_compact"]
    acc = {p: [] for p in PH}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, n_distinct = step_device()
        st = counter.stats()
        for p in PH:
            acc[p].append(st[p])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    launches = ok.launch_count() - launches0
    clocks = sampler.stop()
    st = counter.stats()
    windows, distinct = st["n_windows"], st["n_distinct"]
    phases = {p[3:]: float(np.mean(v)) for p, v in acc.items()}

    # ---- the same step without the capacity hint (the reference's DashMap::new() takes none) ----
    c0 = ok.KmerCounter(K, ok.NORMALIZED, 0)
    for _ in range(2):
        c0.clear(); c0.add_batch_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads); c0.finish_device(1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    nh_steps = max(2, args.steps // 2)
    for _ in range(nh_steps):
        c0.clear(); c0.add_batch_device(d_bases.data_ptr(), n_bases, d_off.data_ptr(), n_reads); c0.finish_device(1)
    torch.cuda.synchronize()
    dt_no_hint = (time.perf_counter() - t0) / nh_steps
    c0.close()

    # ---- e2e: host buffers through the C ABI ------------------------------------------------
    for _ in range(max(1, args.warmup)):
        step_host()
    torch.cuda.synchronize()
    e2e_t["add_batch_ms"].clear(); e2e_t["finish_ms"].clear()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n_out = step_host()
    torch.cuda.synchronize()
    dt_e2e = (time.perf_counter() - t0) / args.steps

    # ---- e2e, two jobs in flight (informational) ---------------------------------------------
    # Two counters driven by two host threads: job i+1's H2D copy runs under job i's D2H copy (PCIe is full
    # duplex; one job alone leaves each direction idle half the time).  Every job still copies its own input
    # and its own result inside the timed region.  Reported beside, never instead of, the sequential e2e.
    dt_pipe = None
    try:
        counter2 = ok.KmerCounter(K, ok.NORMALIZED, hint)
        per_thread = max(2, args.steps)

        def worker(cn, n_jobs):
            for _ in range(n_jobs):
                cn.clear()
                cn.add_batch_ptr(bases.ctypes.data, off.ctypes.data, n_reads)
                pk, pc, _n = cn.finish_raw(1)
                cn.free_result(pk, pc)

        worker(counter2, 1)                       # warm-up: its buffers and page-locked result blocks
        torch.cuda.synchronize()
        th = [threading.Thread(target=worker, args=(cn, per_thread)) for cn in (counter, counter2)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        dt_pipe = (time.perf_counter() - t0) / (2 * per_thread)
        counter2.close()
    except Exception as e:                        # informational only: never fail the bench line on it
        print(f"[bench] pipelined e2e skipped: {e}", file=sys.stderr)

    # ---- roofline of the dominant kernel ----------------------------------------------------
    # Algorithmic bytes per SURVEY.md 8(d): count = B(1+.25+.25) + 16 W, readout = 32 D.  The path
    # that ran decides which kernel dominates: the partitioned path spends its time in the two
    # scatters and the shared-memory count kernel, the table path in k_extract<SinkCount>.
    peak, peak_src = measured_peak()
    alg_count = n_bases * 1.5 + windows * 16.0
    alg_step = alg_count + distinct * 32.0
    if st["partitioned"]:
        cand = {"k_part_scatter_bases (pack+extract+level-1 multisplit)": (phases["scatter1"], n_bases * 1.5 + windows * 8.0),
                "k_part_scatter_keys<2> (level-2 multisplit)": (phases["scatter2"], windows * 16.0),
                "k_part_count (hashed shared-memory dedupe + bucket-ordered emit)": (phases["count"], windows * 8.0 + distinct * 16.0)}
    else:
        cand = {"k_extract<SinkCount> (fused pack+extract+count)": (phases["insert"], alg_count)}
    kname = max(cand, key=lambda n: cand[n][0])
    k_ms, k_bytes = cand[kname]
    achieved = k_bytes / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get(kname.split(" ")[0].split("<")[0])
        except Exception:
            traffic = None

    base, sample = cpu_baseline(n_reads, genome_len, args.sample_reads)
    # parity spot check of the timed configuration: the sample's table through the same library
    sb, so, wk, wc = sample
    chk = ok.KmerCounter(K)
    chk.add_batch(sb, so)
    gk, gc = chk.finish(1)
    chk.close()
    parity_ok = bool(np.array_equal(gk, wk) and np.array_equal(gc, wc))

    line = {
        "metric": METRIC, "value": n_bases / dt, "unit": "bases/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "value_no_hint": n_bases / dt_no_hint, "ms_per_step_no_hint": dt_no_hint * 1e3,
        "capacity_hint": {"value": int(hint), "note": "expected distinct k-mers (0.17 x bases; the truth is 0.143): a speed knob only, "
                                                      "value_no_hint is the same step without it"},
        "timing": {"value_from": "host clock around the K steps, device synchronised on both sides (includes the one host "
                                 "round trip per step)",
                   "ms_per_step_cuda_events": phases["insert"] + phases["readout"],
                   "note": "CUDA events on the library's stream bracket every phase; their sum is the device time of a step"},
        "config": workload_config(n_reads, genome_len, 1),
        "e2e": {"value": n_bases / dt_e2e, "unit": "bases/s", "ms_per_step": dt_e2e * 1e3,
                "h2d_bytes_per_step": int(n_bases + (n_reads + 1) * 8), "d2h_bytes_per_step": int(16 * n_out),
                "add_batch_ms": float(np.mean(e2e_t["add_batch_ms"])), "finish_ms": float(np.mean(e2e_t["finish_ms"])),
                "numa_node": numa_node, "numa_note": NUMA_NOTE["why"],
                "two_jobs_in_flight": None if dt_pipe is None else {
                    "value": n_bases / dt_pipe, "unit": "bases/s", "ms_per_job": dt_pipe * 1e3,
                    "note": "informational: two counters, two host threads, job i+1's H2D under job i's D2H"},
                "note": "add_batch = H2D pieces overlapped with the level-1 scatter; finish = level 2 + count + "
                        "compaction in 16 key-range slices under the D2H of the slices already finished"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                     "algorithmic_bytes_per_launch": k_bytes, "kernel_ms": k_ms,
                     "step": {"algorithmic_bytes": alg_step, "achieved": alg_step / dt / 1e9,
                              "frac": alg_step / dt / 1e9 / peak}},
        "phases_ms": phases,
        "path": "partitioned" if st["partitioned"] else "table",
        "table": {"windows": int(windows), "distinct": int(distinct), "slots": int(st["n_slots"]),
                  "max_displacement": int(st["max_displacement"]), "spilled": int(st["n_spilled"]),
                  "grows": int(st["n_grows"])},
        "cpu_baseline": base,
        "parity_sample_ok": parity_ok,
    }
    print(json.dumps(line))
    counter.close()


# ------------------------------------------------------------------------------ configs[0] --
def run_config1(args, ok, synth, torch, local):
    """one 5 Mbp genome record (10 runs of 100 N, 1 % lower case), k = 21: the whole table against the oracle"""
    import oracle
    oracle.build()
    k = 21
    g = synth.config1_genome()
    off = np.array([0, len(g)], np.uint64)
    bases, _ = pinned_array(ok, len(g), np.uint8)
    bases[:] = g
    d_b = torch.from_numpy(bases).cuda()
    d_o = torch.from_numpy(off.view(np.int64)).cuda()
    counter = ok.KmerCounter(k, ok.NORMALIZED, 0)

    def step_device():
        counter.clear()
        counter.add_batch_device(d_b.data_ptr(), len(g), d_o.data_ptr(), 1)
        return counter.finish_device(1)

    def step_host():
        counter.clear()
        counter.add_batch_ptr(bases.ctypes.data, off.ctypes.data, 1)
        pk, pc, n = counter.finish_raw(1)
        counter.free_result(pk, pc)
        return n

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ok.launch_count()
    steps = max(args.steps, 20)                       # a step is ~1 ms: a few more for a stable clock sample
    t0 = time.perf_counter()
    for _ in range(steps):
        step_device()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    launches = (ok.launch_count() - launches0) // steps
    clocks = sampler.stop()
    st = counter.stats()
    for _ in range(2):
        step_host()
    t0 = time.perf_counter()
    for _ in range(steps):
        n_out = step_host()
    dt_e2e = (time.perf_counter() - t0) / steps
    keys, counts = counter.finish(1)
    t0 = time.perf_counter()
    wk, wc = oracle.count_batch(k, g, off)
    t_cpu = time.perf_counter() - t0
    parity = bool(np.array_equal(keys, wk) and np.array_equal(counts, wc))
    text_ok = ok.format_counts(keys[:1000], counts[:1000], k) == oracle.format_counts(wk[:1000], wc[:1000], k)
    peak, peak_src = measured_peak()
    alg = len(g) * 1.5 + st["n_windows"] * 16.0 + st["n_distinct"] * 32.0
    line = {"metric": "bases/sec counted (canonical k=21)", "value": len(g) / dt, "unit": "bases/s", "n_gpus": 1, "steps": steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": dict(CONFIG1, bases=int(len(g))),
            "e2e": {"value": len(g) / dt_e2e, "unit": "bases/s", "ms_per_step": dt_e2e * 1e3, "h2d_bytes_per_step": int(len(g) + 16),
                    "d2h_bytes_per_step": int(16 * n_out)},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole step (a 5 Mbp genome is ~1 ms of launches; no kernel dominates)",
                         "achieved": alg / dt / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / dt / 1e9 / peak, "peak_source": peak_src,
                         "traffic": None, "algorithmic_bytes_per_launch": alg},
            "phases_ms": {p[3:]: float(st[p]) for p in ("ms_sample", "ms_scatter1", "ms_scatter2", "ms_count", "ms_compact", "ms_insert", "ms_readout")},
            "path": "partitioned" if st["partitioned"] else "table",
            "table": {"windows": int(st["n_windows"]), "distinct": int(st["n_distinct"])},
            "cpu_baseline": {"value": len(g) / t_cpu, "unit": "bases/s", "cores": 1, "kind": "port",
                             "sample": f"the whole genome ({len(g)} bases), {t_cpu:.2f} s", "host_cores_available": os.cpu_count()},
            "parity_full_table_ok": parity, "parity_tsv_text_ok": bool(text_ok)}
    print(json.dumps(line))
    counter.close()


# ----------------------------------------------------- checkers of the set configurations (oracle) --
def oracle_compare_sample(k, n_sets, length, sizes, inter, n_pairs):
    """a few pairs of the all-vs-all against the oracle's compare (compare.rs:51-66); also the timed CPU baseline
    (1 thread: the reference compares two databases per process, single-threaded)"""
    import oracle
    import bench_sets
    from orion_kmer_b200 import synth
    oracle.build()
    cand = [(0, 1), (0, min(49, n_sets - 1)), (0, min(50, n_sets - 1)), (n_sets // 3, n_sets - 1), (n_sets // 2, n_sets // 2 + 1)]
    pairs = [p for p in dict.fromkeys(cand) if p[0] != p[1]][:max(1, n_pairs)]
    need = sorted({i for p in pairs for i in p})
    osets = {i: oracle.kmer_set_batch(k, bench_sets.genome(synth, i, length), np.array([0, length], np.uint64)) for i in need}
    ok_all, t = True, 0.0
    for i, j in pairs:
        t0 = time.perf_counter()
        r = oracle.compare(osets[i], osets[j])
        t += time.perf_counter() - t0
        ok_all &= (r["db1"], r["db2"], r["intersection_size"]) == (int(sizes[i]), int(sizes[j]), int(inter[i, j]))
        ok_all &= int(inter[j, i]) == int(inter[i, j]) and int(inter[i, i]) == int(sizes[i])
    return {"ok": bool(ok_all), "pairs": [list(p) for p in pairs],
            "cpu_baseline": {"value": len(pairs) / t, "unit": "pairs/s", "cores": 1, "kind": "port",
                             "sample": f"{len(pairs)} pairs of {length}-base genome sets (hash one set, probe with the other), {t:.2f} s; "
                                       "set construction not included", "host_cores_available": os.cpu_count()}}


def oracle_query_sample(k, n_genomes, length, reads, roff, hits, n_reads, cpu_genomes):
    """Parity of the build + union + query path at a size the CPU can check: the SAME library calls on a subset of
    cpu_genomes genomes and n_reads reads against the oracle (build.rs:46-58, db_types.rs:43-48, query.rs:83-107);
    properties of the full-size result; the timed CPU baseline (build: 1 thread; query: all cores, query.rs:78)."""
    import oracle
    import bench_sets
    import orion_kmer_b200 as ok
    from orion_kmer_b200 import synth
    oracle.build()
    subset = sorted({int(x) for x in np.linspace(0, n_genomes - 1, min(cpu_genomes, n_genomes))})
    one = np.array([0, length], np.uint64)
    t0 = time.perf_counter()
    first = oracle.kmer_set_batch(k, bench_sets.genome(synth, subset[0], length), one)          # faithful single thread
    t_build1 = time.perf_counter() - t0
    nt = min(os.cpu_count() or 1, 64)
    osets = [first] + [oracle.count_batch_ranged_mt(k, bench_sets.genome(synth, i, length), one, nt)[0] for i in subset[1:]]
    ounion = np.unique(np.concatenate(osets))
    r_n = min(n_reads, len(roff) - 1)
    rb, ro = reads[:r_n * 150], roff[:r_n + 1]
    t0 = time.perf_counter()
    want = oracle.query_hits(ounion, k, rb, ro, nt)
    t_query = time.perf_counter() - t0
    gsets = []
    for i in subset:
        s = ok.KmerSet.build(k)
        s.add_batch(bench_sets.genome(synth, i, length), one)
        gsets.append(s)
    gunion = ok.KmerSet.union(gsets)
    got = gunion.probe_reads(rb, ro, ok.RAW)
    same = bool(len(gunion) == len(ounion) and np.array_equal(got.astype(np.uint64), want)
                and all(len(s) == len(o) for s, o in zip(gsets, osets)))
    for s in gsets + [gunion]:
        s.close()
    # full-size properties: no read has more hits than windows; adding genomes can only add hits
    props = bool(int(hits.max()) <= 150 - k + 1 and np.all(hits[:r_n].astype(np.uint64) >= want))
    return {"ok": bool(same and props), "reads": int(r_n), "subset_genomes": len(subset),
            "cpu_baseline": {"value": length / t_build1, "unit": "bases/s", "cores": 1, "kind": "port",
                             "sample": f"build of 1 genome ({length} bases, 1 thread, {t_build1:.2f} s); query of {r_n} reads against the union of "
                                       f"{len(subset)} genomes on {nt} threads: {r_n / t_query:.0f} reads/s",
                             "query_reads_per_s": r_n / t_query, "query_cores": nt, "host_cores_available": os.cpu_count()}}


def run_sets_config(args, torch, ok, synth, world, rank, local):
    import bench_sets
    ctx = {"ok": ok, "synth": synth, "torch": torch, "world": world, "rank": rank, "local": local, "ClockSampler": ClockSampler,
           "measured_peak": measured_peak, "oracle_compare_sample": oracle_compare_sample, "oracle_query_sample": oracle_query_sample,
           "pinned": lambda n: pinned_array(ok, n, np.uint8)[0]}
    line = bench_sets.run_compare(args, ctx) if args.config == 5 else bench_sets.run_build_query(args, ctx)
    bench_sets.emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def run_reference_sets(args):
    """--impl reference for configs 4 / 5: the oracle port on a bounded sample, all the host threads the reference
    itself would use (query.rs:78 is its only parallel loop)"""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n = args.sets if args.config == 5 else args.genomes
    vals = []
    for _ in range(args.warmup + args.steps):
        if args.config == 5:
            sizes = {}
            chk = None
            import oracle
            import bench_sets
            from orion_kmer_b200 import synth
            oracle.build()
            a = oracle.kmer_set_batch(21, bench_sets.genome(synth, 0, args.genome_len), np.array([0, args.genome_len], np.uint64))
            b = oracle.kmer_set_batch(21, bench_sets.genome(synth, 1, args.genome_len), np.array([0, args.genome_len], np.uint64))
            t0 = time.perf_counter()
            for _ in range(3):
                oracle.compare(a, b)
            base = {"value": 3 / (time.perf_counter() - t0), "unit": "pairs/s", "cores": 1, "kind": "port",
                    "sample": "3 compares of one pair of 5 Mbp genome sets"}
        else:
            import oracle
            import bench_sets
            from orion_kmer_b200 import synth
            oracle.build()
            g = bench_sets.genome(synth, 0, args.genome_len)
            t0 = time.perf_counter()
            oracle.kmer_set_batch(31, g, np.array([0, len(g)], np.uint64))
            base = {"value": len(g) / (time.perf_counter() - t0), "unit": "bases/s", "cores": 1, "kind": "port",
                    "sample": "build of 1 genome, single thread (build.rs:93-116 is sequential)"}
        vals.append(base["value"])
    v = float(np.mean(vals[args.warmup:]))
    base["value"] = v
    metric = ("set pairs compared per second (k=21 all-vs-all: |A|, |B|, |A n B| per pair; Jaccard on the host)" if args.config == 5
              else "bases/sec through build + union + query (k=31)")
    print(json.dumps({"impl": "reference", "metric": metric, "value": v, "unit": base["unit"], "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                      "config": {"workload": f"BASELINE.json configs[{args.config - 1}] ({n} sets)"}, "cpu_baseline": base,
                      "e2e": {"value": v, "unit": base["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU")
    ap.add_argument("--genome", type=int, default=0, help="genome length (default 5 x reads = 30x coverage)")
    ap.add_argument("--hint", type=int, default=0, help="expected distinct k-mers per GPU")
    ap.add_argument("--sample-reads", type=int, default=250_000,
                    help="reads in the CPU baseline sample (250k = 37.5 M bases: ~10-20 s of single-thread CPU work)")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json configs[config-1]")
    ap.add_argument("--total-reads", type=int, default=0, help="config 3: reads of the whole job, split over the GPUs (default 100 M)")
    ap.add_argument("--sets", type=int, default=256, help="config 5: genome sets")
    ap.add_argument("--genomes", type=int, default=1000, help="config 4: genomes in the database")
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--query-reads", type=int, default=1_000_000, help="config 4: reads in the query sample")
    ap.add_argument("--parity-pairs", type=int, default=5)
    ap.add_argument("--parity-reads", type=int, default=20_000)
    ap.add_argument("--cpu-genomes", type=int, default=12, help="config 4: genomes in the CPU-checked subset")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the parity check of the sharded table")
    ap.add_argument("--parity-seconds", type=float, default=60.0, help="N > 1: CPU budget of the key-slice check")
    args = ap.parse_args()
    if args.config == 3:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        args.total_reads = args.total_reads or 100_000_000
        args.reads = args.total_reads // max(1, world)
        args.genome = args.genome or args.total_reads * 5 // max(1, world)      # multi.bench multiplies by world: 30x coverage of the whole job
    if args.impl == "reference":
        if args.config in (4, 5):
            run_reference_sets(args)
        elif args.config == 1:
            run_reference_config1(args)
        else:
            run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
