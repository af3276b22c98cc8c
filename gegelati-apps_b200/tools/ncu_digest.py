#!/usr/bin/env python3
"""Digest of an ncu capture for profiles/: key raw metrics, executed-SASS opcode mix and the hottest
CUDA source lines.  usage: ncu_digest.py <base>   (expects <base>_raw.csv, <base>_src.csv, <base>_cs.csv made with
  ncu -i X.ncu-rep --page raw --csv ; --page source --csv ; --page source --csv --print-source cuda,sass)"""
import collections
import csv
import re
import sys

base = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
r = list(csv.reader(open(base + "_raw.csv")))
h, u, v = r[0], r[1], r[2]
print("== raw metrics (%s)" % v[h.index("Kernel Name")] if "Kernel Name" in h else "== raw metrics")
for k in KEYS:
    if k in h:
        print("%-72s %s %s" % (k, v[h.index(k)], u[h.index(k)]))
for i, k in enumerate(h):
    if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(v[i] or 0) > 0.05:
        print("%-72s %s" % (k.replace("smsp__average_warps_issue_stalled_", "stall/issue: ").replace("_per_issue_active.ratio", ""), v[i]))

rows = list(csv.reader(open(base + "_src.csv")))
hdr = rows[1]
ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot, by = 0, collections.Counter()
for row in rows[2:]:
    if len(row) <= max(ia, isrc) or not row[ia].isdigit():   # a second kernel's header / partial rows
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", row[isrc].strip())
    op = m.group(2).split(".")[0] if m else row[isrc]
    by[op] += int(row[ia]); tot += int(row[ia])
print("\n== executed warp instructions by opcode (total %d, %d SASS lines)" % (tot, len(rows) - 2))
print("  ".join("%s %.1f%%" % (op, 100.0 * n / tot) for op, n in by.most_common(28)))

rows = list(csv.reader(open(base + "_cs.csv")))
cur, hdr, agg = None, None, []
for row in rows:
    if len(row) >= 2 and row[0] == "File Path":
        cur = row[1].split("/")[-1]; continue
    if row and row[0] == "Line No":
        hdr = row; ia, ismp = row.index("Instructions Executed"), row.index("# Samples"); continue
    if hdr and len(row) > max(ia, ismp) and row[0] != "" and row[0].isdigit() and row[ia].isdigit():
        agg.append((cur, int(row[0]), row[1].strip()[:100], int(row[ia]), int(row[ismp]) if row[ismp].isdigit() else 0))
merged = collections.OrderedDict()   # the same line can appear once per captured launch / inlined copy
for f, l, src, n, sm in agg:
    k = (f, l, src)
    merged[k] = (merged.get(k, (0, 0))[0] + n, merged.get(k, (0, 0))[1] + sm)
agg = [(k[0], k[1], k[2], v[0], v[1]) for k, v in merged.items()]
ti, ts = sum(a[3] for a in agg), sum(a[4] for a in agg)
print("\n== hottest source lines (share of executed instructions | share of stall samples)")
for f, l, s, n, sm in sorted(agg, key=lambda a: -a[3])[:top]:
    print("%5.2f%% %5.2f%%  %s:%d  %s" % (100.0 * n / ti, 100.0 * sm / ts, f, l, s))
