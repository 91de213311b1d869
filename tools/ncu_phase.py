#!/usr/bin/env python3
"""Attribute an ncu report's samples to the enclosing top-level source line (inlined helper code is charged to the last
line of `anchor_file` seen in SASS address order): tools/ncu_phase.py <file.ncu-rep> <anchor_file> [bucket]"""
import csv, subprocess, sys, io, collections
rep, anchor = sys.argv[1], sys.argv[2]; bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 10
import os
KSEL = (["--kernel-name", os.environ["NCU_KERNEL"]] if os.environ.get("NCU_KERNEL") else [])
txt = subprocess.run(["ncu", "-i", rep] + KSEL + ["--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
insts = []; cur_file = None; H = None; cur_line = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": H = r; ai = H.index("Address"); si = H.index("# Samples"); ii = H.index("Instructions Executed"); continue
    if H is None or len(r) <= ii: continue
    if r[0].strip(): cur_line = int(r[0])
    if not r[ai].strip(): continue
    try: insts.append((int(r[ai], 16), cur_file, cur_line, int(r[si] or 0), int(r[ii] or 0)))
    except ValueError: pass
insts.sort(); agg = collections.OrderedDict(); last = 0
for addr, f, ln, s, i in insts:
    if f == anchor: last = (ln // bucket) * bucket
    agg.setdefault(last, [0, 0, 0]); agg[last][0] += s; agg[last][1] += i; agg[last][2] += 1
tot = sum(v[0] for v in agg.values()) or 1; toti = sum(v[1] for v in agg.values()) or 1
print("total samples %d, warp instructions %d, static instructions %d" % (tot, toti, len(insts)))
for k, (s, i, n) in sorted(agg.items()):
    print("%s:%4d-%4d  %5.1f%% smp %5.1f%% inst  %5d static" % (anchor, k, k + bucket - 1, 100.0 * s / tot, 100.0 * i / toti, n))
