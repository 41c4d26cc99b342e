#!/usr/bin/env python
"""Turns an ncu launch list of one bench.py step (CSV written by

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --quick --batch B

) into profiles/<name>_launches.md (per-kernel totals of the LAST step: launches, time, share of the step, DRAM bytes) and
profiles/r2_traffic.json (dram bytes per launch of the igemm / wgrad families, keyed by the build hash; bench.py reads
`roofline.traffic` from it and withholds it when the build differs).

usage: python scripts/ncu_traffic.py gpurun_out/launches.csv <workload> <per_gpu_batch> <out_md> [--json profiles/r2_traffic.json]
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def family(name):
    if "igemm" in name:
        return "igemm"
    if "wgrad_kernel" in name and "hwgrad" not in name and "edge" not in name:
        return "wgrad"
    return None


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("fmri::", "")
    return re.sub(r"\(.*", "", name)[:70]


def main():
    path, workload, batch, out_md = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    jpath = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    iid, iname, imetric, iunit, ival = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(int(r[iid]), dict(name=r[iname]))
        v = float(r[ival].replace(",", ""))
        unit = r[iunit]
        if r[imetric].startswith("gpu__time_duration"):
            d["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
        else:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
            d[r[imetric].split(".")[0]] = v * mult
    seq = list(launches.values())
    # the capture holds warm-up + timed step(s): keep the last one (the sequence of kernel names repeats)
    names = [d["name"] for d in seq]
    n = len(names)
    # every step launches vgan_gate_kernel exactly once, followed by the same optimizer + repack tail: the last step starts
    # right after the previous step's tail (steps are not launch-for-launch identical: lazily created buffers, fills)
    gates = [i for i, nm in enumerate(names) if "vgan_gate" in nm]
    step = n
    if len(gates) >= 2:
        tail = n - 1 - gates[-1]
        step = n - (gates[-2] + tail + 1)
    last = seq[-step:]
    tot_ns = sum(d.get("ns", 0) for d in last)
    agg = OrderedDict()
    for d in last:
        a = agg.setdefault(short(d["name"]), dict(n=0, ns=0.0, rd=0.0, wr=0.0))
        a["n"] += 1
        a["ns"] += d.get("ns", 0)
        a["rd"] += d.get("dram__bytes_read", 0)
        a["wr"] += d.get("dram__bytes_write", 0)
    with open(out_md, "w") as f:
        f.write(f"# ncu launch list of one {workload} step at batch {batch} per GPU (the last {step} of {n} captured launches)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` -- per-launch "
                "times are cold-cache and serialised: the SHARE of the step is what compares with bench.py's CUDA-event numbers.\n\n")
        f.write(f"{step} launches, {tot_ns / 1e6:.2f} ms summed kernel time, DRAM {sum(a['rd'] + a['wr'] for a in agg.values()) / 1e9:.1f} GB\n\n")
        f.write("| kernel | launches | ms | share | DRAM read GB | DRAM write GB | GB/s |\n|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda t: -t[1]["ns"]):
            gbs = (a["rd"] + a["wr"]) / max(a["ns"], 1)
            f.write(f"| {k} | {a['n']} | {a['ns'] / 1e6:.3f} | {100 * a['ns'] / tot_ns:.1f} % | {a['rd'] / 1e9:.2f} | {a['wr'] / 1e9:.2f} | {gbs:.0f} |\n")
    if jpath:
        import bench

        fams = {}
        for d in last:
            fam = family(d["name"])
            if fam:
                a = fams.setdefault(fam, dict(launches=0, bytes=0.0, ns=0.0))
                a["launches"] += 1
                a["bytes"] += d.get("dram__bytes_read", 0) + d.get("dram__bytes_write", 0)
                a["ns"] += d.get("ns", 0)
        out = dict(build_hash=bench.build_hash(), workload=workload, per_gpu_batch=batch, source=os.path.basename(out_md),
                   families={k: dict(launches=v["launches"], dram_bytes_per_launch=v["bytes"] / v["launches"],
                                     ms_total_under_ncu=v["ns"] / 1e6, share_of_step=v["ns"] / tot_ns) for k, v in fams.items()})
        with open(jpath, "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
