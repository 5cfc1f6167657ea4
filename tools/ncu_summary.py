#!/usr/bin/env python3
"""Summarise an .ncu-rep: key raw metrics + per-opcode executed/stall shares from the source page.
usage: tools/ncu_summary.py report.ncu-rep [kernel-substring]"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
pat = re.compile(r"^(dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|smsp__inst_executed\.sum|sm__inst_executed_pipe_\w+\.sum|"
                 r"sm__pipe_\w+_cycles_active\.avg\.pct_of_peak_sustained_active|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                 r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__occupancy_limit_\w+|"
                 r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|smsp__cycles_active\.avg|sm__cycles_elapsed\.avg|"
                 r"sm__icc_request_hit_rate\.pct|l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|"
                 r"sm__inst_executed_pipe_\w+\.avg\.pct_of_peak_sustained_active|smsp__inst_executed_pipe_\w+\.sum|launch__grid_size|launch__block_size)$")
for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    if len(sys.argv) > 2 and sys.argv[2] not in name:
        continue
    print("== kernel:", name[:100])
    for h, u, v in zip(hdr, units, vals):
        if pat.match(h) and v not in ("", "0"):
            print(f"  {h:75s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# one section per captured launch, each starting with its own header row
heads = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r]
for n, hi in enumerate(heads[:3]):                      # the first three launches are enough for a summary
    hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
    end = heads[n + 1] if n + 1 < len(heads) else len(rows)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ex, sm, st = collections.Counter(), collections.Counter(), collections.Counter()
    for r in rows[hi + 1:end]:
        if len(r) < len(hdr): continue
        t = r[ix["Source"]].strip()
        if not t or not (r[ix["Instructions Executed"]] or "0").isdigit(): continue
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        ex[op] += int(r[ix["Instructions Executed"]] or 0); sm[op] += int(r[ix["# Samples"]] or 0)
        for h in stalls: st[h] += int(r[ix[h]] or 0)
    te, ts = sum(ex.values()), sum(sm.values())
    print(f"== source page (launch {n}): {te} warp instructions, {ts} samples")
    for op, e in ex.most_common(18):
        print(f"  {op:20s} exec {100 * e / max(1, te):5.1f}%   samples {100 * sm[op] / max(1, ts):5.1f}%")
    print("  stalls:", {k[6:]: round(100 * v / max(1, ts), 1) for k, v in st.most_common(8)})
