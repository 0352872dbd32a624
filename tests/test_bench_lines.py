"""The JSON lines of bench.py that can be produced without a GPU: the reference arm of every configuration (the oracle
port on a bounded sample) and -- with a numpy stand-in for the device library -- the all-vs-all line of bench_sets.py.
Checked: the keys the bench contract names, the same `config` object on both arms, a roofline fraction that follows
from the stated bytes and time."""
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "vs_baseline", "dtype", "data", "config",
             "cpu_baseline", "e2e"}


def _reference(*args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *args],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout                      # ONE JSON line
    return json.loads(lines[0])


@pytest.mark.parametrize("config,extra", [(2, ["--sample-reads", "3000"]), (1, []), (4, ["--genome-len", "200000"]), (5, ["--genome-len", "200000"])])
def test_reference_arm_lines(config, extra):
    d = _reference("--config", str(config), *extra)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_under_torchrun_env_prints_on_rank_0_only_and_matches_our_config():
    env = {"WORLD_SIZE": "8", "RANK": "0"}
    d = _reference("--config", "3", "--gpus", "8", "--sample-reads", "2000", env=env)
    assert d["scaling"] == "strong" and d["n_gpus"] == 8
    sys.path.insert(0, ROOT)
    import bench
    from orion_kmer_b200 import multi
    ours = multi.bench_config(bench.workload_config, 3, 12_500_000, 500_000_000, 8)        # what multi.bench puts into our line
    assert d["config"] == ours and ours["total_reads"] == 100_000_000 and ours["sub_batches_per_rank"] == 2
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "3", "--gpus", "8"],
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, WORLD_SIZE="8", RANK="5"))
    assert r.returncode == 0 and r.stdout.strip() == ""                                    # the other ranks exit 0 without work
    w = _reference("--config", "2", "--gpus", "4", "--sample-reads", "2000", env={"WORLD_SIZE": "4", "RANK": "0"})
    assert w["config"] == multi.bench_config(bench.workload_config, 2, 10_000_000, 200_000_000, 4) and w["scaling"] == "weak"


class _FakeSet:
    def __init__(self, keys):
        self.keys = np.ascontiguousarray(keys, dtype=np.uint64)

    def __len__(self):
        return len(self.keys)

    def to_array(self):
        return self.keys

    def close(self):
        pass


def test_all_vs_all_line_with_a_stand_in_device(oracle):
    """bench_sets.run_compare end to end, the device library replaced by the oracle + numpy (the parity leg inside the
    bench then compares the oracle with itself: what is tested is the line, not the kernels)"""
    sys.path.insert(0, ROOT)
    import bench
    import bench_sets
    from orion_kmer_b200 import synth
    launches = {"n": 0}

    class KmerSet:
        @staticmethod
        def build(k):
            s = _FakeSet(np.zeros(0, np.uint64))
            s.k = k
            s.add_batch = lambda b, o: setattr(s, "keys", oracle.kmer_set_batch(k, b, o))
            return s

        @staticmethod
        def from_sorted(k, a):
            return _FakeSet(a)

    def all_vs_all(sets):
        launches["n"] += 4
        n = len(sets)
        inter = np.zeros((n, n), np.uint64)
        for i in range(n):
            for j in range(n):
                inter[i, j] = len(np.intersect1d(sets[i].keys, sets[j].keys, assume_unique=True))
        return np.array([len(s) for s in sets], np.uint64), inter

    fake_ok = types.SimpleNamespace(KmerSet=KmerSet, all_vs_all=all_vs_all, launch_count=lambda: launches["n"])
    fake_torch = types.SimpleNamespace(cuda=types.SimpleNamespace(synchronize=lambda: None))
    sampler = type("S", (), {"__init__": lambda self, gpu: None, "start": lambda self: None,
                             "stop": lambda self: {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}})
    args = types.SimpleNamespace(sets=6, genome_len=30_000, steps=2, warmup=1, parity_pairs=3)
    ctx = {"ok": fake_ok, "synth": synth, "torch": fake_torch, "world": 1, "rank": 0, "local": 0, "ClockSampler": sampler,
           "measured_peak": lambda: (6535.4, "test"), "oracle_compare_sample": bench.oracle_compare_sample}
    d = bench_sets.run_compare(args, ctx)
    json.dumps(d)                                            # serialisable
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks", "ms_per_step", "scaling"} <= set(d)
    assert d["parity_pairs_ok"] is True and d["config"]["pairs"] == 15 and d["n_gpus"] == 1
    rf = d["roofline"]
    assert rf["kernel"].startswith("k_ava_tiles") and set(rf) >= {"bound", "achieved", "peak", "unit", "frac", "traffic", "pairwise_model"}
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert abs(rf["achieved"] - rf["algorithmic_bytes_per_launch"] / (d["ms_per_step"] / 1e3) / 1e9) < 1e-6 * max(1.0, rf["achieved"])
    assert rf["algorithmic_bytes_per_launch"] < rf["pairwise_model"]["bytes"]            # every key once < every set once per pair
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * d["config"]["keys_total"] and d["e2e"]["d2h_bytes_per_step"] == 6 * 6 * 8
