#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here with `ncu -i <rep> --page raw --csv`) into a small table:
per captured launch: grid, duration, DRAM bytes read+written, DRAM / tensor-pipe utilisation, registers.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    for i, h in enumerate(hdr):
        if h == name or h.endswith("." + name) or h.endswith(name):
            return i
    return None


want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_%")]
idx = [(col(n), lab, n) for n, lab in want]
print(f"# ncu --set full summary of `{rep}`\n")
print("| " + " | ".join(lab for i, lab, n in idx if i is not None) + " |")
print("|" + "---|" * sum(1 for i, _, _ in idx if i is not None))
for r in data:
    cells = []
    for i, lab, n in idx:
        if i is None:
            continue
        v = r[i]
        if lab == "kernel":
            v = v.split("(")[0].replace("void fmri::", "")
        u = units[i]
        cells.append(f"{v} {u}".strip())
    print("| " + " | ".join(cells) + " |")
missing = [n for i, lab, n in idx if i is None]
if missing:
    print("\nmetrics not present in this report: " + ", ".join(missing))
