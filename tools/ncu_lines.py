"""Per-CUDA-line summary of an ncu report's source page (stall samples, instructions executed).

usage: python tools/ncu_lines.py report.ncu-rep [--top 40] [--file k_msckf.cu] [--ranges 100-200,201-300]
Reads `ncu -i report --page source --csv --print-source cuda,sass` (the report must have been captured with
--import-source on and the library built with -lineinfo)."""
import argparse
import csv
import io
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, hdr, data = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "" or not r[0].isdigit():
            continue
        d = dict(zip(hdr[4:], r[4:]))
        try:
            data.append((cur_file, int(r[0]), r[1], int(d["# Samples"]), int(d["Instructions Executed"]), d))
        except (KeyError, ValueError):
            pass
    return data


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--file", default=None)
    ap.add_argument("--ranges", default=None)
    a = ap.parse_args()
    data = load(a.rep)
    tot_s = sum(d[3] for d in data) or 1
    tot_i = sum(d[4] for d in data) or 1
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    byfile = {}
    for f, ln, src, s, i, _ in data:
        byfile.setdefault(f, [0, 0])
        byfile[f][0] += s
        byfile[f][1] += i
    for f, (s, i) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f}: samples {100 * s / tot_s:.1f}%  instr {100 * i / tot_i:.1f}%")
    sel = [d for d in data if a.file is None or d[0].endswith(a.file)]
    if a.ranges:
        for rg in a.ranges.split(","):
            lo, hi = (int(x) for x in rg.split("-"))
            s = sum(d[3] for d in sel if lo <= d[1] <= hi)
            i = sum(d[4] for d in sel if lo <= d[1] <= hi)
            print(f"lines {lo}-{hi}: samples {100 * s / tot_s:.1f}%  instr {100 * i / tot_i:.1f}%")
    for f, ln, src, s, i, d in sorted(sel, key=lambda d: -d[3])[: a.top]:
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "(" not in k and v.isdigit() and int(v) > 0}
        top = ",".join(f"{k}:{v}" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:3])
        print(f"{f.split('/')[-1]}:{ln:4d} {100 * s / tot_s:5.1f}% instr {100 * i / tot_i:5.1f}%  [{top}]  {src.strip()[:90]}")


if __name__ == "__main__":
    sys.exit(main())
