#!/usr/bin/env python
"""Per-source-line stall-sample histogram from an ncu report captured with --import-source on.
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n] [file_filter]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
flt = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = None; hdr = None; lines = collections.OrderedDict(); total = 0
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] and r[0].isdigit() and hdr:
        try: smp = int(r[4]); ins = int(r[7])
        except Exception: continue
        key = (cur_file.split('/')[-1], int(r[0]))
        if key in lines: lines[key] = (lines[key][0] + smp, lines[key][1] + ins, r[1])
        else: lines[key] = (smp, ins, r[1])
        total += smp
print("total samples", total)
items = [(k, v) for k, v in lines.items() if not flt or flt in k[0]]
for k, v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:4d} {100.0*v[0]/max(total,1):5.1f}%  inst {v[1]:>10d}  {v[2].strip()[:110]}")
# coarse buckets of 20 lines for the main file
if len(sys.argv) > 4:
    # stage buckets "name:lo-hi,name:lo-hi" over the filtered file
    tot = collections.OrderedDict()
    for spec in sys.argv[4].split(","):
        nm, rg = spec.split(":"); lo, hi = map(int, rg.split("-"))
        s = sum(v[0] for k, v in items if lo <= k[1] <= hi); i = sum(v[1] for k, v in items if lo <= k[1] <= hi)
        print(f"  {nm:12s} {100.0*s/max(total,1):5.1f}%  inst {i}")
    other = sum(v[0] for k, v in lines.items() if flt and flt not in k[0])
    print(f"  other files  {100.0*other/max(total,1):5.1f}%")
