#!/usr/bin/env python3
"""Key raw metrics of EVERY kernel in one ncu report + DRAM traffic per robot: tools/ncu_kernels.py <file.ncu-rep> --batch B [--json traffic.json]"""
import csv, subprocess, sys, io, json
args = sys.argv[1:]; out_json = None; batch = None
if "--json" in args: i = args.index("--json"); out_json = args[i + 1]; del args[i:i + 2]
if "--batch" in args: i = args.index("--batch"); batch = int(args[i + 1]); del args[i:i + 2]
rep = args[0]
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
NAMES = {"mpc_flow_kernel": "flow", "mpc_lq_kernel": "lq", "mpc_riccati_kernel": "riccati", "mpc_linesearch_kernel": "linesearch", "wbc_update_kernel": "wbc"}
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt))); hdr, units = rows[0], rows[1]; summary = {}
for r in rows[2:]:
    kname = r[hdr.index("Kernel Name")]; short = next((v for k, v in NAMES.items() if k in kname), kname[:24])
    print("== %s  (%s)\n   %s" % (short, rep.split("/")[-1], kname[:110])); vals = {}
    for i, h in enumerate(hdr):
        if h in KEYS: print("   %-68s %-16s %s" % (h, units[i], r[i])); vals[h] = (r[i], units[i])
        elif "issue_stalled" in h and "per_issue_active" in h:
            try:
                if float(r[i]) > 0.2: print("   stall %-62s %s" % (h.split("issue_stalled_")[1].split("_per_issue")[0], r[i]))
            except ValueError: pass
    rd = float(vals["dram__bytes_read.sum"][0]) * UNIT[vals["dram__bytes_read.sum"][1]]; wr = float(vals["dram__bytes_write.sum"][0]) * UNIT[vals["dram__bytes_write.sum"][1]]
    tu = vals["gpu__time_duration.sum"][1]
    summary[short] = {"report": rep.split("/")[-1], "batch": batch, "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_robot": (rd + wr) / batch if batch else None,
                      "kernel_ms": float(vals["gpu__time_duration.sum"][0]) * (1.0 if tu == "ms" else 1e-3 if tu == "us" else 1e3)}
if out_json: json.dump(summary, open(out_json, "w"), indent=1)
