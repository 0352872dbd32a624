"""`orion-kmer-b200` without a GPU: everything the command line decides BEFORE any device work -- clap-style usage
errors (exit code 2), k validation with the texts of errors.rs (exit code 1, count.rs:43-45, build.rs:83-85,
classify.rs:69-72) -- and the loud failure of every command that would need the device ("no CPU fallback", exit
code 1, no output file).  The same binary runs the reference's integration tests in tests/test_gpu_cli.py."""
import os
import subprocess

import pytest

import orion_kmer_b200 as ok


@pytest.fixture(scope="module")
def exe():
    ok.build_host()
    p = ok.cli_path()
    assert os.path.exists(p), "orion-kmer-b200 is not built (ok.build_host())"
    return p


def run(exe, *args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")          # no device, also on a GPU box
    return subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=120, env=env)


def test_help_lists_the_reference_commands_and_flags(exe):
    r = run(exe, "--help")
    assert r.returncode == 0
    for word in ("count", "build", "compare", "query", "classify", "--threads", "--verbose"):      # cli.rs:5-35
        assert word in r.stdout
    for sub, flags in (("count", ("-k", "-i", "-o", "-m")), ("build", ("-k", "-g", "-o")), ("compare", ("--db1", "--db2", "-o")),
                       ("query", ("-d", "-r", "-o", "-c")), ("classify", ("-i", "-d", "-o", "--min-kmer-frequency"))):
        line = next(ln for ln in r.stdout.splitlines() if ln.startswith(sub + ":"))
        assert all(f in line for f in flags), line


@pytest.mark.parametrize("args", [
    ("count", "-k", 21, "-o", "o.tsv"),                       # cli.rs:43 input files are required
    ("count", "-i", "a.fa", "-o", "o.tsv"),                   # cli.rs:40 -k is required
    ("count", "-k", 300, "-i", "a.fa", "-o", "o.tsv"),        # -k is a u8
    ("count", "-k", "x", "-i", "a.fa", "-o", "o.tsv"),
    ("build", "-k", 21, "-o", "o.db"),                        # cli.rs:68
    ("compare", "--db1", "a.db", "-o", "o.json"),             # cli.rs:85
    ("query", "-d", "a.db", "-o", "o.txt"),                   # cli.rs:108
    ("classify", "-i", "a.fa", "-o", "o.json"),               # cli.rs:143
    ("bogus",),
    (),
])
def test_usage_errors_exit_with_2_like_clap(exe, args, tmp_path):
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=120, cwd=tmp_path)
    assert r.returncode == 2, (args, r.stderr)
    assert r.stderr.strip() and not os.listdir(tmp_path)


@pytest.mark.parametrize("k", [0, 33, 255])
def test_invalid_k_is_reported_before_any_device_work(exe, k, tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_text(">s\nACGTACGT\n")
    for args in (("count", "-k", k, "-i", fa, "-o", tmp_path / "o.tsv"),          # count.rs:43-45, count_tests.rs:296-331
                 ("build", "-k", k, "-g", fa, "-o", tmp_path / "o.db"),           # build.rs:83-85
                 ("classify", "-i", fa, "-d", tmp_path / "none.db", "-o", tmp_path / "o.json", "-k", k)):   # classify.rs:69-72
        r = run(exe, *args)
        assert r.returncode == 1, (args, r.stderr)
        assert f"Invalid K-mer size: {k}. Must be between 1 and 32." in r.stderr          # errors.rs:6-7
    assert sorted(os.listdir(tmp_path)) == ["a.fa"]


def test_without_a_device_every_command_fails_loudly(exe, tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_text(">s\nACGTACGTACGTAAAC\n")
    db = tmp_path / "a.db"
    ok.write_kmer_db(str(db), 4, {"a.fa": [1, 2, 3]})
    for args in (("count", "-k", 4, "-i", fa, "-o", tmp_path / "o.tsv"),
                 ("build", "-k", 4, "-g", fa, "-o", tmp_path / "o.db"),
                 ("compare", "--db1", db, "--db2", db, "-o", tmp_path / "o.json"),
                 ("query", "-d", db, "-r", fa, "-o", tmp_path / "o.txt"),
                 ("classify", "-i", fa, "-d", db, "-o", tmp_path / "o2.json")):
        r = run(exe, *args)
        assert r.returncode == 1, (args, r.stdout, r.stderr)
        assert "no CPU fallback" in r.stderr, (args, r.stderr)
    assert sorted(os.listdir(tmp_path)) == ["a.db", "a.fa"]           # nothing was written
