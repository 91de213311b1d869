#!/usr/bin/env python3
"""Summarise an ncu report's source page per CUDA source line: tools/ncu_hot.py <file.ncu-rep> [top_n]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
import os
KSEL = (["--kernel-name", os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else [])
txt = subprocess.run(["ncu", "-i", rep] + KSEL + ["--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
agg = collections.OrderedDict(); cur_file = None; H = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": H = r; si = H.index("# Samples"); ii = H.index("Instructions Executed"); continue
    if H is None or len(r) <= ii: continue
    try: smp = int(r[si] or 0); ins = int(r[ii] or 0)
    except ValueError: continue
    if r[0].strip():
        key = (cur_file, r[0], r[1].strip()[:130]); agg.setdefault(key, [0, 0]); cur = key
    else:
        key = cur
    agg[key][0] += smp; agg[key][1] += ins
tot = sum(v[0] for v in agg.values()) or 1; toti = sum(v[1] for v in agg.values()) or 1
print("total samples %d, instructions %d" % (tot, toti))
for (f, ln, src), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%s  %s" % (100.0 * s / tot, 100.0 * i / toti, f, ln, src))
