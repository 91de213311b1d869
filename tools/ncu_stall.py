#!/usr/bin/env python3
"""Per CUDA source line: samples of one stall reason (ncu source page): NCU_KERNEL=name tools/ncu_stall.py <file.ncu-rep> <stall column, e.g. stall_long_sb> [top_n]"""
import csv, subprocess, sys, io, collections, os
rep, col = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
KSEL = (["--kernel-name", os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else [])
txt = subprocess.run(["ncu", "-i", rep] + KSEL + ["--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt))); agg = collections.OrderedDict(); cur_file = None; H = None; cur = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": H = r; si = H.index(col); ti = H.index("# Samples"); continue
    if H is None or len(r) <= si: continue
    try: smp = int(r[si] or 0); tot = int(r[ti] or 0)
    except ValueError: continue
    if r[0].strip(): cur = (cur_file, r[0], r[1].strip()[:150]); agg.setdefault(cur, [0, 0])
    if cur is None: continue
    agg[cur][0] += smp; agg[cur][1] += tot
allc = sum(v[0] for v in agg.values()) or 1; alls = sum(v[1] for v in agg.values()) or 1
print("%s: %d of %d samples (%.1f%%)" % (col, allc, alls, 100.0 * allc / alls))
for (f, ln, src), (s, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%%  %s:%s  %s" % (100.0 * s / allc, f, ln, src))
