#!/usr/bin/env python
"""Markdown results table from the bench JSON lines of a round (profiles/r2_bench_*.json written by scripts/final_benches.sh).
usage: python scripts/make_results_table.py profiles/r2_bench_*.json"""
import json
import os
import sys

rows = []
for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    except Exception as e:  # noqa: BLE001
        rows.append((os.path.basename(path), f"unreadable: {e}", "", "", "", ""))
        continue
    r = d.get("roofline") or {}
    cpu = d.get("cpu_baseline") or {}
    st = d.get("stock_torch") or {}
    cfg = d.get("config", {})
    name = os.path.basename(path).replace("r2_bench_", "").replace(".json", "")
    rows.append((name,
                 f"{d['value']:,.0f} ({d['ms_per_step']:.2f} ms)" + (f"; e2e {d['e2e']['value']:,.0f}" if d.get("e2e") and d.get("impl") != "reference" else ""),
                 f"{100 * r['whole_step_frac']:.1f} %" if r.get("whole_step_frac") else "-",
                 f"{r['achieved']:.0f} TF/s = {r['frac']:.2f}" if r.get("achieved") else "-",
                 f"{st['value']:,.0f}" if st.get("value") else "-",
                 f"{cpu['value']:.1f} ({cpu.get('kind')}, {cpu.get('cores')} cores)" if cpu.get("value") else "-",
                 f"{(d.get('clocks') or {}).get('sm_mhz', '-')}" + (" graph" if cfg.get("cuda_graph") else "")))
print("| run | samples/s (ms/step) | whole step, % of 1380 TF/s | dominant kernel family | stock PyTorch, same GPU | host-CPU reference | SM MHz |")
print("|---|---|---|---|---|---|---|")
for r in rows:
    print("| " + " | ".join(r) + " |")
