#!/usr/bin/env python3
"""ncu csv of tools/collect_r03.sh step (2) -> profiles/r03_riccati_dmma.json (read by bench.py): tools/dmma_json.py <dmma.csv> <batch> <regular_nodes> <out.json>"""
import csv, json, sys
allrows = list(csv.reader(open(sys.argv[1]))); h0 = next(i for i, r in enumerate(allrows) if r and r[0] == "ID")   # the bench line and ncu banners precede the table
hdr = allrows[h0]; vals = {}
for r in [r for r in allrows[h0 + 1:] if len(r) == len(hdr)]:
    d = dict(zip(hdr, r)); vals[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
batch, nodes = int(sys.argv[2]), int(sys.argv[3])
out = {"dmma_per_node": vals["sm__inst_executed_pipe_tensor_subpipe_dmma.sum"] / (batch * nodes), "tensor_pipe_active_pct": vals["smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"],
       "dmma_instructions": vals["sm__inst_executed_pipe_tensor_subpipe_dmma.sum"], "batch": batch, "regular_nodes": nodes, "kernel_time": vals.get("gpu__time_duration.sum"), "source": "profiles/" + sys.argv[1].split("/")[-1]}
json.dump(out, open(sys.argv[4], "w"), indent=1); print(out)
