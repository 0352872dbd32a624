#!/usr/bin/env python
"""SASS listings of the hot kernels as shipped in liborion_gpu.so -> profiles/sass_<kernel>.txt + sass_summary.txt.

Runs without a GPU (cuobjdump reads the sm_100a cubin inside the library).  `python tools/sass_listing.py`
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "orion_kmer_b200", "liborion_gpu.so")

# short name -> regular expression the mangled function name must match (first match wins)
KERNELS = [
    ("k_part_scatter_bases_k31_p3", r"^_Z20k_part_scatter_basesILb0ELb0ELi31ELb1EE"),
    ("k_part_scatter_keys_level2_k31", r"^_Z19k_part_scatter_keysILi2ELb1ELi31ELb0EE"),
    ("k_part_scatter_keys_level1_strided", r"^_Z19k_part_scatter_keysILi1ELb0ELi0ELb1EE"),
    ("k_part_count_13_k31", r"^_Z12k_part_countILi13ELi31EE"),
    ("k_part_count_generic", r"^_Z20k_part_count_generic"),
    ("k_merge_write", r"^_Z13k_merge_writeILb1EE"),
    ("k_intersect_row_tiled", r"^_Z21k_intersect_row_tiled"),
    ("k_member_tiled", r"^_Z14k_member_tiled"),
    ("k_ava_bounds", r"^_Z12k_ava_bounds"),
    ("k_ava_tiles", r"^_Z11k_ava_tiles"),
    ("k_xchg_sample", r"^_Z13k_xchg_sample"),
]
# mnemonics worth counting: TMA bulk copies, mbarriers, atomics, votes, barriers, tensor-core ops (none expected)
MARKS = ["UBLKCP", "SYNCS", "ATOMS", "ATOMG", "REDG", "MATCH", "VOTE", "BAR", "CCTL", "POPC", "UTCMMA", "LDTM", "HMMA"]


def functions():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    cur, body = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            body[cur].append((m.group(1), m.group(2).strip()))
    return body


def main():
    body = functions()
    summary = ["# SASS of the hot kernels as shipped in liborion_gpu.so (sm_100a cubin, nvcc 12.9); full listings: profiles/sass_<kernel>.txt",
               "# UBLKCP = 1-D TMA bulk copy, SYNCS = mbarrier ops, ATOMS = shared-memory atomics; no UTCMMA / LDTM / HMMA: nothing here is a contraction",
               "# regenerate: python tools/sass_listing.py (no GPU needed)"]
    for short, pat in KERNELS:
        name = next((f for f in body if re.match(pat, f)), None)
        if name is None:
            print("no function matches", pat, file=sys.stderr)
            continue
        ins = body[name]
        ops = collections.Counter()
        for _, text in ins:
            t = re.sub(r"^@!?U?P\d+\s+", "", text)
            ops[t.split()[0].split(".")[0]] += 1
        hist = ", ".join(f"{o} {c}" for o, c in ops.most_common(14))
        with open(os.path.join(ROOT, "profiles", f"sass_{short}.txt"), "w") as f:
            f.write(f"# cuobjdump -sass liborion_gpu.so (sm_100a), function {name}\n")
            f.write(f"# {len(ins)} instructions; opcode histogram: {hist}\n")
            for addr, text in ins:
                f.write(f"/*{addr}*/ {text}\n")
        marks = " ".join(f"{m}:{ops[m]}" for m in MARKS if ops[m])
        summary.append(f"{short:38s} {len(ins):5d} instr  {marks}")
    with open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w") as f:
        f.write("\n".join(summary) + "\n")
    print("\n".join(summary))


if __name__ == "__main__":
    main()
