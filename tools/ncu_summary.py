#!/usr/bin/env python3
"""Key metrics of an .ncu-rep (raw page) and, with --source, the hottest source lines by stall
samples.  usage: ncu_summary.py report.ncu-rep [--source N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__cycles_active.avg"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
for r in rows[2:]:
    print("KERNEL", r[h.index("Kernel Name")][:80], "id", r[h.index("ID")])
    for k in KEYS:
        if k in h:
            i = h.index(k)
            print(f"   {k} [{u[i]}] = {r[i]}")
    for i, n in enumerate(h):
        if "warp_issue_stalled" in n and n.endswith("_per_warp_active.pct"):
            try:
                if float(r[i]) > 3:
                    print(f"   {n.replace('smsp__average_warps_issue_stalled_','stall ').replace('_per_warp_active.pct','')} = {r[i]}")
            except ValueError:
                pass
if "--source" in sys.argv:
    n = int(sys.argv[sys.argv.index("--source") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    fname, hh, data = "", None, []
    for r in rows:
        if r and r[0] == "File Name":
            fname = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hh = r
        elif hh and r and r[0].isdigit():
            try:
                data.append((float(r[hh.index("# Samples")]), float(r[hh.index("Instructions Executed")]), fname, r[0], r[1]))
            except (ValueError, IndexError):
                pass
    tot = sum(d[0] for d in data) or 1
    toti = sum(d[1] for d in data) or 1
    print(f"--- top {n} source lines by stall samples (total samples {tot:.0f}, warp instructions {toti:.3e})")
    for s_, i_, f_, l_, t_ in sorted(data, reverse=True)[:n]:
        print(f"{100 * s_ / tot:5.1f}% smp {100 * i_ / toti:5.1f}% inst  {f_}:{l_:>4}  {t_.strip()[:100]}")
