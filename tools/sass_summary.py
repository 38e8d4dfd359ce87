#!/usr/bin/env python
"""Per-kernel SASS evidence for profiles/: which Blackwell/Hopper-era mnemonics each kernel of libgnn_b200.so
contains (B200_PROFILING.md "What proves a Blackwell-native kernel").  Runs without a GPU:
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
Counts per kernel: UBLKCP (cp.async.bulk, the 1-D TMA engine: .S.G = global->shared gather, .G.S = shared->global
store), SYNCS (mbarrier), LDG.E.128 / STG.E.128 (128-bit vector access), FADD2 (packed fp32 add), UTMALDG/UTMASTG
(tensor-map TMA), UTC*MMA / HMMA (tensor cores: none expected — aggregation is bandwidth-bound), ATOM/RED (atomics:
only the integer counters of the wave mover and CUB's radix sort are expected), registers and static shared memory."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "graphneuralnetwork_b200", "libgnn_b200.so")
PATTERNS = [("UBLKCP.S.G", r"\bUBLKCP\.S\.G"), ("UBLKCP.G.S", r"\bUBLKCP\.G\.S"), ("SYNCS", r"\bSYNCS\."),
            ("LDG.128", r"\bLDG\.E\.128"), ("STG.128", r"\bSTG\.E\.128"), ("LDG", r"\bLDG\."), ("STG", r"\bSTG\."),
            ("FADD2", r"\bFADD2\b"), ("FFMA", r"\bFFMA\b"), ("SHFL", r"\bSHFL\."), ("MUFU.EX2", r"\bMUFU\.EX2"),
            ("UTMA*", r"\bUTMA(LDG|STG)"), ("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("HMMA", r"\bHMMA\b"),
            ("ATOM/RED", r"\b(ATOM|ATOMG|RED)\."), ("MEMBAR.SYS", r"\bMEMBAR\.[A-Z.]*SYS"), ("ST.sys", r"\bSTG\.E[A-Z.]*\.SYS")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            regs[cur] = tuple(int(x) for x in m.groups())
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS:
            if re.search(pat, line):
                counts[cur][name] += 1
    names = demangle(list(counts))
    ours = [k for k in counts if "cub::" not in names[k]]
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (sm_100a): {len(ours)} kernels of this repo, "
          f"{len(counts) - len(ours)} of CUB (radix sort / scan of the CSR builders)\n")
    tot = collections.Counter()
    for k in counts:
        tot.update(counts[k])
    print("totals over all kernels: " + ", ".join(f"{n} x{tot[n]}" for n, _ in PATTERNS) + "\n")
    short = lambda s: re.sub(r"\(anonymous namespace\)::|<unnamed>::|gnn::", "", s)[:150]
    keys = ["UBLKCP.S.G", "UBLKCP.G.S", "SYNCS", "LDG.128", "STG.128", "FADD2", "SHFL", "MUFU.EX2", "ATOM/RED", "UTC*MMA", "HMMA"]
    print("kernel | regs | stack | " + " | ".join(keys))
    for k in ours:
        if not (counts[k]["UBLKCP.S.G"] or counts[k]["LDG.128"] or counts[k]["ATOM/RED"] or counts[k]["SHFL"] > 40):
            continue  # the small helper kernels are listed in the totals only
        r = regs.get(k, ("?", "?", "?"))
        print(f"{short(names[k])} | {r[0]} | {r[1]} | " + " | ".join(str(counts[k][n]) for n in keys))


if __name__ == "__main__":
    sys.exit(main())
