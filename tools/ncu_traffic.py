#!/usr/bin/env python
"""profiles/r02_traffic.json from raw ncu CSV pages: per kernel of interest, dram__bytes_read.sum +
dram__bytes_write.sum per launch (what bench.py reports as `roofline.traffic`), duration, registers, DRAM %.
    python tools/ncu_traffic.py profiles/r02_prof_sage_raw.csv [more.csv ...]
The bench's gather launch covers both hops + the sources' own rows in ONE launch of sage_tma_kernel: key "sage_multi"."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write", "gpu__time_duration.sum": "duration",
        "launch__registers_per_thread": "registers", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct",
        "smsp__inst_executed.sum": "warp_instructions"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def parse(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        rec = {"kernel": r[hdr.index("Kernel Name")]}
        for k, name in WANT.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    rec[name] = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
                except ValueError:
                    pass
        out.append(rec)
    return out


def main():
    res = {}
    for path in sys.argv[1:]:
        for rec in parse(path):
            k = rec["kernel"]
            key = None
            if "sage_tma_kernel" in k:
                key = "sage_multi"
            elif "gat_fwd_kernel" in k:
                key = "gat_fwd_bf16" if "bfloat16" in k else "gat_fwd"
            elif "gat_bwd2_kernel" in k:
                targs = k.split("gat_bwd2_kernel<")[1].split(">")[0].replace("(bool)", "").replace("(int)", "").split(",")
                tr = targs[4].strip() in ("1", "true")
                key = "gat_bwd_pass2_transposed" if tr else "gat_bwd_pass1_rows"
            elif "spmm_rbs_kernel" in k:
                key = "spmm_rbs"
            if key is None:
                continue
            rec["dram_bytes"] = rec.get("dram_read", 0.0) + rec.get("dram_write", 0.0)
            rec["source"] = os.path.relpath(path, ROOT)
            if key not in res or rec["dram_bytes"] > res[key]["dram_bytes"]:
                res[key] = rec  # several launches of a kernel: keep the largest (the bench's full-size launch)
    out = os.path.join(ROOT, "profiles", "r02_traffic.json")
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
