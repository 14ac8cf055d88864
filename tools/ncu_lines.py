#!/usr/bin/env python
"""Per-source-line hot spots from an ncu report (needs -lineinfo and --import-source on).
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n] [launch_index]
"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = None; hdr = None
agg = collections.OrderedDict(); seen_files = set(); second = False
for r in csv.reader(raw.splitlines()):
    if not r: continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        if cur_file in seen_files: second = True     # files repeat once per captured launch
        seen_files.add(cur_file)
        continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if second: continue                                   # first captured launch only
    if hdr is None or r[0] == "": continue
    d = dict(zip(hdr[4:], r[4:]))
    key = (cur_file, int(r[0]), r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0, 0, 0])
    num = lambda k: int(d[k]) if d.get(k, "").isdigit() else 0
    a[0] += num("# Samples"); a[1] += num("Instructions Executed")
    a[2] += num("L1 Wavefronts Shared"); a[3] += num("L1 Wavefronts Shared Excessive")
ts = sum(a[0] for a in agg.values()) or 1; ti = sum(a[1] for a in agg.values()) or 1; tw = sum(a[2] for a in agg.values()) or 1
print(f"total samples {ts}  warp-instructions {ti}  shared wavefronts {tw}")
print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s} {'smem%':>6s} {'excess':>9s}  source")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0][:22]+':'+str(k[1]):28s} {a[0]/ts*100:6.2f} {a[1]/ti*100:6.2f} {a[2]/tw*100:6.2f} {a[3]:9d}  {k[2]}")
