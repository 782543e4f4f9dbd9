#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): key metrics + hottest SASS lines.  usage: ncu_summary.py rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_bytes.sum"]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for wn in want:
    if wn in hdr:
        i = hdr.index(wn); print(f"{wn} [{units[i]}]: {[r[i] for r in rows[2:]]}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None: cur["rows"].append(r)
b = blocks[0]; hdr = b["rows"][0]; data = b["rows"][1:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ils, iss = hdr.index("stall_long_sb"), hdr.index("stall_short_sb")
tot = sum(int(r[isamp] or 0) for r in data); totex = sum(int(r[iex] or 0) for r in data)
print("kernel", b["name"][:70], "samples", tot, "warp-instructions", totex, "sass lines", len(data))
print("samples  executed  long_sb short_sb  sass")
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:ntop]:
    print(f"{r[isamp]:>6} {r[iex]:>9} {r[ils]:>6} {r[iss]:>6}  {r[isrc][:100]}")
