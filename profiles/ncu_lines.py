#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line:
   python profiles/ncu_lines.py export.csv [top_n]   -> samples and executed instructions per file:line"""
import csv
import sys
import collections

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_samp = hdr.index("# Samples")
        i_inst = hdr.index("Instructions Executed")
        continue
    if hdr is None or r[0] in ("Function Name",):
        continue
    if r[0] != "":  # a source line row (aggregate of its SASS)
        key = (cur_file, int(r[0]), r[1].strip()[:90])
        try:
            s = int(r[i_samp]) if r[i_samp] not in ("-", "") else 0
        except ValueError:
            s = 0
        agg.setdefault(key, [0, 0])
        agg[key][0] += s
    else:
        if r[2] == "...":
            continue
        try:
            agg[key][1] += int(r[i_inst])
        except (ValueError, NameError):
            pass
tot_s = sum(v[0] for v in agg.values()) or 1
tot_i = sum(v[1] for v in agg.values()) or 1
print("total samples %d, total warp-instructions %d" % (tot_s, tot_i))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% samp %5.1f%% inst  %s:%d  %s" % (100.0 * v[0] / tot_s, 100.0 * v[1] / tot_i, k[0], k[1], k[2]))
