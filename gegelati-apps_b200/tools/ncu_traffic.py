#!/usr/bin/env python3
"""profiles/k1_traffic.json from an ncu capture of K1 (what bench.py quotes as roofline.traffic / roofline.ncu):
DRAM bytes per launch, executed FP64 warp instructions and the fraction of the FP64 issue rate they are.
usage: ncu_traffic.py <base> <digest name>   (expects <base>_raw.csv and <base>_cs.csv as made for ncu_digest.py)"""
import csv
import json
import re
import sys

base, digest = sys.argv[1], sys.argv[2]
r = list(csv.reader(open(base + "_raw.csv")))
h, v = r[0], r[2]


def m(name):
    return float(v[h.index(name)].replace(",", ""))


rows = list(csv.reader(open(base + "_cs.csv")))
hdr = next(i for i, row in enumerate(rows) if row and row[0] == "Line No")
H = rows[hdr]
isrc, iex = [i for i, x in enumerate(H) if x == "Source"][1], H.index("Instructions Executed")
fp64 = 0
for row in rows[hdr + 1:]:
    if len(row) > iex and row[iex]:
        op = row[isrc].strip()
        op = op.split()[1] if op.startswith("@") and len(op.split()) > 1 else (op.split()[0] if op else "")
        if re.match(r"(DADD|DMUL|DFMA)\b", op):
            fp64 += int(float(row[iex]))
unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
ub = r[1]


def dram(name):
    return m(name) * unit.get(ub[h.index(name)], 1.0)


t_ms = m("gpu__time_duration.sum")
sm, clk = m("launch__grid_size"), 1.965e9
out = {
    "kernel": "tpgx_bid2_kernel<true> (K1 v2, team mode)",
    "workload": "C2: 2000 roots x 5 edges x 20 lines, 60 000 samples (tools/profile_case.py 2000 60000 1)",
    "dram_bytes_per_launch": dram("dram__bytes_read.sum") + dram("dram__bytes_write.sum"),
    "gpu_time_ms": t_ms,
    "fp64_pipe_pct": m("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": m("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warp_instructions": m("smsp__inst_executed.sum"),
    "executed_fp64_warp_instructions": fp64,
    "executed_fp64_flop_frac_of_peak": fp64 * 32.0 / (t_ms * 1e-3) / (148 * 64 * clk),
    "executed_fp64_note": "DADD + DMUL + DFMA thread instructions per cycle / (64 per SM per cycle at 1.965 GHz); a DFMA counted as ONE instruction; DSETP not counted",
    "l1_hit_pct": m("l1tex__t_sector_hit_rate.pct"),
    "l2_hit_pct": m("lts__t_sector_hit_rate.pct"),
    "shared_wavefronts": m("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    "source": "profiles/%s (ncu --set full --clock-control none, one launch)" % digest,
}
print(json.dumps(out, indent=1))
