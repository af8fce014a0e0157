#!/usr/bin/env python
"""Per-region view of an ncu SASS source page: instructions executed and stall samples,
grouped into runs of N consecutive SASS instructions (default 64), with the dominant stall
reasons. Usage: ncu -i rep --page source --print-source sass --csv > x.csv; ncu_sass_hot.py x.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hdr = rows[1]; data = rows[2:]
ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot_i = sum(int(r[ci] or 0) for r in data); tot_s = sum(int(r[cs] or 0) for r in data)
print("total inst %d samples %d" % (tot_i, tot_s))
tot_st = {h: sum(int(r[i] or 0) for r in data) for i, h in stall_cols}
print("stalls:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / tot_s) for h, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:8]))
for b in range(0, len(data), N):
    blk = data[b:b + N]
    ins = sum(int(r[ci] or 0) for r in blk); smp = sum(int(r[cs] or 0) for r in blk)
    st = {h: sum(int(r[i] or 0) for r in blk) for i, h in stall_cols}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    ops = {}
    for r in blk:
        op = r[1].split()[0] if not r[1].startswith("@") else r[1].split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
    topops = ",".join("%s%d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:4])
    print("%5d %5.1f%% inst %5.1f%% smp | %-44s | %s" % (b, 100.0 * ins / tot_i, 100.0 * smp / tot_s,
          " ".join("%s %.0f%%" % (h[6:], 100.0 * v / max(smp, 1)) for h, v in top), topops))
