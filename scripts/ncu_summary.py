#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here with `ncu -i <rep> --page raw --csv`) into a small table:
per captured launch: grid, duration, DRAM bytes read+written, DRAM / tensor-pipe utilisation, registers.
usage: python scripts/ncu_summary.py gpurun_out/a.ncu-rep|a.csv [more ...] > profiles/<name>.md
"""
import csv
import subprocess
import sys

reps = sys.argv[1:]
raws = []
for rep in reps:   # .ncu-rep reports, or the `--page raw --csv` export of one (scripts/ncu_capture.sh keeps only that)
    raws.append(open(rep).read() if rep.endswith(".csv") else
                subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
rep = ", ".join(reps)
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("sm__inst_executed.sum", "warp_inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_%")]
print(f"# ncu --set full summary ({len(reps)} captures, one launch each unless noted)\n")
print("| " + " | ".join(lab for _, lab in want) + " |")
print("|" + "---|" * len(want))
for raw in raws:      # every report has its own (auto-scaled) unit row
    rows = list(csv.reader([l for l in raw.splitlines() if l.startswith('"')]))
    if len(rows) < 3:
        continue
    hdr, units, data = rows[0], rows[1], rows[2:]

    def col(name):
        for i, h in enumerate(hdr):
            if h == name or h.endswith("." + name) or h.endswith(name):
                return i
        return None

    for r in data:
        cells = []
        for n, lab in want:
            i = col(n)
            if i is None:
                cells.append("-")
                continue
            v = r[i]
            if lab == "kernel":
                v = v.split("(fmri")[0].replace("void fmri::", "").replace("fmri::", "").replace("(int)", "")[:60]
            else:
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
            cells.append(f"{v} {units[i]}".strip())
        print("| " + " | ".join(cells) + " |")
