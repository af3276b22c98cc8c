#!/usr/bin/env python3
"""Per-handler view of a K1 v2 capture: splits the SASS of the interpreter loop at the branches back to the fetch
and prints executed warp instructions, executions (= executions of the segment's last instruction) and stall samples.
usage: ncu_sass_segments.py <base>_cs.csv   (ncu -i X.ncu-rep --page source --csv --print-source cuda,sass)"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
H = rows[hdr]
ia, isrc, iex, ist = H.index("Address"), [i for i, h in enumerate(H) if h == "Source"][1], H.index("Instructions Executed"), H.index("# Samples")
ilsb, issb = H.index("stall_long_sb"), H.index("stall_short_sb")
ins = []
for r in rows[hdr + 1:]:
    if len(r) <= iex or not r[ia]:
        continue
    try:
        ins.append((int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia]), r[isrc].strip(), int(float(r[iex] or 0)), int(float(r[ist] or 0)),
                    int(float(r[ilsb] or 0)), int(float(r[issb] or 0))))
    except ValueError:
        pass
ins.sort()
total = sum(i[2] for i in ins)
tot_s = sum(i[3] for i in ins)
brx = [k for k, i in enumerate(ins) if i[1].startswith("BRX")]
print("total executed warp instructions %.4g, stall samples %d, BRX at %s" % (total, tot_s, [hex(ins[k][0]) for k in brx]))
k0 = brx[0]
# fetch = the LDS.128 a few instructions before the first BRX
f = max(k for k in range(k0 - 10, k0) if ins[k][1].startswith("LDS.128"))
fetch_addr = ins[f][0]
disp = ins[f:k0 + 1]
print("dispatch: %d instructions, executed %.4g each (= entries), %.2f%% of all instructions, %.1f%% of stall samples" %
      (len(disp), disp[-1][2], 100.0 * sum(i[2] for i in disp) / total, 100.0 * sum(i[3] for i in disp) / tot_s))
seg, cur = [], []
end = None
for k in range(k0 + 1, len(ins)):
    cur.append(ins[k])
    m = re.match(r"(@!?U?P\d+\s+)?BRA(\.U)?\s+0x([0-9a-f]+)", ins[k][1])
    if m and int(m.group(3), 16) == fetch_addr and not m.group(1):
        seg.append(cur)
        cur = []
    if ins[k][1].startswith("BSYNC") and not seg == [] and cur and len(cur) > 0 and k > k0 + 800:
        break
loop_instr = sum(len(s) for s in seg)
print("handler segments: %d, %d SASS instructions (%.1f KB)" % (len(seg), loop_instr, loop_instr * 16 / 1024.0))
inloop = sum(i[2] for s in seg for i in s) + sum(i[2] for i in disp)
print("share of executed instructions inside the interpreter loop: %.1f%%" % (100.0 * inloop / total))
print("%5s %12s %8s %7s %7s %7s  %s" % ("size", "executed", "share%", "stall%", "longsb", "shortsb", "first / fp64 ops"))
for s in seg:
    ex = sum(i[2] for i in s)
    st = sum(i[3] for i in s)
    ops = [i[1].split()[0] if not i[1].startswith("@") else i[1].split()[1] for i in s]
    fp = sum(1 for o in ops if o.startswith(("DADD", "DMUL", "DFMA", "DSETP")))
    print("%5d %12.4g %8.2f %7.2f %7d %7d  %s fp64=%d" % (len(s), ex, 100.0 * ex / total, 100.0 * st / tot_s, sum(i[4] for i in s), sum(i[5] for i in s),
                                                     " ".join(ops[:3]), fp))
