#!/usr/bin/env python
"""SASS-level stall breakdown of the first captured launch of an ncu report.
    python tools/ncu_sass.py gpurun_out/prof.ncu-rep [first_row] [n_rows]
Prints per instruction: samples, executed count, top stall reasons.  Also a per-opcode-class summary."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; out = []; kernels = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1: break
        continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or not r: continue
    out.append(dict(zip(hdr, r)))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(d["# Samples"] or 0) for d in out)
print("instructions", len(out), "samples", tot)
agg = collections.Counter(); aggst = collections.defaultdict(collections.Counter); cnt = collections.Counter()
for d in out:
    op = d["Source"].split()[0] if not d["Source"].strip().startswith("@") else d["Source"].split()[1]
    op = op.split(".")[0]
    s = int(d["# Samples"] or 0); agg[op] += s; cnt[op] += int(d["Instructions Executed"] or 0)
    for k in stalls: aggst[op][k] += int(d[k] or 0)
print("\nper opcode: samples%  executed  top stalls")
for op, s in agg.most_common(25):
    top = ", ".join(f"{k[6:]}={v}" for k, v in aggst[op].most_common(4) if v)
    print(f"{op:10s} {s / tot * 100:6.2f}%  {cnt[op]:>12d}  {top}")
allst = collections.Counter()
for op in aggst:
    allst.update(aggst[op])
print("\nall stalls:", ", ".join(f"{k[6:]}={v / tot * 100:.1f}%" for k, v in allst.most_common(10)))
if len(sys.argv) > 2:
    a = int(sys.argv[2]); n = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    for i, d in enumerate(out[a:a + n]):
        top = ", ".join(f"{k[6:]}={d[k]}" for k in stalls if int(d[k] or 0) > 0)
        print(f"{a + i:5d} {int(d['# Samples'] or 0):5d} {int(d['Instructions Executed'] or 0):9d}  {d['Source'][:70]:70s} {top}")
