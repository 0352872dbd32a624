"""The C-ABI library loads and exports every symbol include/orion_gpu.h declares.  CPU only:
no compute entry is called except to check that it refuses to run without a device."""
import ctypes
import os
import re

import pytest

import orion_kmer_b200 as ok

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "orion_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ok_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_we_bind():
    assert header_functions() == sorted(ok.ABI)


def test_library_exports_every_declared_symbol():
    ok._build.build()
    L = ctypes.CDLL(ok.gpu_library_path())
    for name in header_functions():
        assert hasattr(L, name), name


def test_version_and_error_strings():
    assert b"sm_100a" in ok.lib().ok_version()
    with pytest.raises(ok.InvalidKmerSize) as e:
        ok.KmerCounter(0)
    assert str(e.value) == "Invalid K-mer size: 0. Must be between 1 and 32."   # errors.rs:6-7
    with pytest.raises(ok.InvalidKmerSize) as e:
        ok.KmerCounter(33)
    assert str(e.value) == "Invalid K-mer size: 33. Must be between 1 and 32."


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry fails loudly (OK_ERR_NO_DEVICE)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ok.OrionError) as e:
        ok.KmerCounter(21)
    assert e.value.code == ok.OK_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_integration_md_binds_only_what_the_header_declares_with_the_same_arity():
    """the Rust `extern "C"` block of INTEGRATION.md (what a maintainer of the reference crate would paste into
    src/gpu.rs) must name functions of include/orion_gpu.h and give each the number of arguments the header gives it"""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = doc[doc.index('extern "C" {'):doc.index("pub const OK_NORM_NORMALIZED")]
    block = re.sub(r"//[^\n]*", "", block)
    rust = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (ok_[a-z0-9_]+)\s*\((.*?)\)\s*->", block, flags=re.S)}
    assert len(rust) >= 20
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "orion_gpu.h")).read(), flags=re.S)
    c = {m.group(1): m.group(2) for m in re.finditer(r"\b(ok_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", header, flags=re.S)}

    def arity(args):
        args = args.strip()
        return 0 if args in ("", "void") else args.count(",") + 1

    for name, args in rust.items():
        assert name in c, f"{name} is bound in INTEGRATION.md but not declared in the header"
        assert arity(args) == arity(c[name]), (name, args, c[name])
