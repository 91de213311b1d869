#!/usr/bin/env python
"""Worst per-block relative error of every GPU parity test: tools/summarize_parity.py <QMB_PARITY_LOG jsonl> > profiles/<tag>_parity_levels.txt

The log is written by tests/_parity.py when QMB_PARITY_LOG is set (one line per comparison: test label + relative error of every block of like quantities)."""
import json, sys, collections

worst = collections.OrderedDict()
count = collections.Counter()
for line in open(sys.argv[1]):
    d = json.loads(line)
    w = worst.setdefault(d["test"], {})
    count[d["test"]] += 1
    for k, v in d["levels"].items():
        w[k] = max(w.get(k, 0.0), float(v))
print("# worst relative error per block of like quantities, CUDA path vs oracle (contract 1e-5; asserted tolerances in tests/_parity.py)")
print("# %-58s %5s  %s" % ("comparison", "n", "block: worst"))
for t, w in worst.items():
    print("%-60s %5d  %s" % (t[:60], count[t], "  ".join("%s %.1e" % (k, v) for k, v in w.items())))
allv = [v for w in worst.values() for v in w.values()]
print("# overall worst: %.2e over %d comparisons" % (max(allv), sum(count.values())))
