#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key roofline metrics + per-opcode lane efficiency.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:70s} {units[i]:12s} {vals[i]}")
for i, h in enumerate(hdr):
    if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h.lower():
        try:
            if float(vals[i]) > 0.2:
                print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:28s} {float(vals[i]):.2f}")
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
byop = collections.defaultdict(lambda: [0, 0, 0])
ti = tt = ts = 0
for r in rows[2:]:
    try:
        ie, te, sm = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]])
    except (ValueError, IndexError):
        continue
    m = re.match(r"\s*(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    op = m.group(1).split(".")[0] if m else "?"
    byop[op][0] += ie; byop[op][1] += te; byop[op][2] += sm
    ti += ie; tt += te; ts += sm
print(f"warp instr {ti:.3e}  thread instr {tt:.3e}  avg active lanes {tt / max(ti, 1):.2f}")
for op, v in sorted(byop.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"  {op:8s} inst {100 * v[0] / ti:5.1f}%  lanes {v[1] / max(v[0], 1):5.1f}  samples {100 * v[2] / max(ts, 1):5.1f}%")
