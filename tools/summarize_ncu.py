#!/usr/bin/env python
"""Turn the two ncu outputs of tools/gpu_round.sh into the text summaries kept under profiles/.

    python tools/summarize_ncu.py gpurun_out/launches.csv gpurun_out/prof.ncu-rep profiles/r01_v1
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = None
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, rows = r, rows[i + 1:]
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# {len(rows)} launches, total {tot:.1f} us\n")
        f.write(f"{'kernel':72s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}  grid block\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:72]:72s} {v[0]:8d} {v[1]:10.1f} {v[1] / v[0]:9.1f} {v[1] / tot * 100:6.1f}%  {v[2]} {v[3]}\n")


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on  (selected raw metrics per captured launch)\n")
        for d in data:
            f.write(f"\n== {d[hdr.index('Kernel Name')]}  grid {d[hdr.index('Grid Size')]} block {d[hdr.index('Block Size')]}\n")
            for k in KEEP:
                if k in hdr:
                    f.write(f"{k:90s} {d[hdr.index(k)]:>18s} {units[hdr.index(k)]}\n")


def traffic_digest(rep, tag, windows_per_launch, path="profiles/ncu_traffic.json"):
    """dram bytes of the largest captured launch of each kernel -> profiles/ncu_traffic.json (read by bench.py)."""
    import json, os
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    def val(d, k):
        i = hdr.index(k)
        v = float(d[i].replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[i]]
    digest = json.load(open(path)) if os.path.exists(path) else {}
    best = {}
    for d in data:
        name = d[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
        tot = val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum")
        def num(k):
            return float(d[hdr.index(k)].replace(",", "")) if k in hdr and d[hdr.index(k)] not in ("", "n/a") else None
        if name not in best or tot > best[name][0]:
            best[name] = (tot, val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum"),
                          num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), num("smsp__inst_executed.sum"),
                          num("sm__cycles_elapsed.max"), num("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                          num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                          num("smsp__issue_active.avg.pct_of_peak_sustained_active"))
    for name, (tot, rd, wr, wav, inst, cyc, fp64, tens, issue) in best.items():
        digest[name] = {"dram_bytes_per_launch": tot, "dram_bytes_read": rd, "dram_bytes_write": wr,
                        "windows_per_launch": windows_per_launch, "dram_bytes_per_window": tot / windows_per_launch,
                        "smem_wavefronts_per_window": wav / windows_per_launch if wav else None,
                        "warp_instructions_per_window": inst / windows_per_launch if inst else None,
                        "sm_cycles_elapsed": cyc, "fp64_pipe_pct": fp64, "tensor_pipe_pct": tens, "issue_active_pct": issue,
                        "source": f"{tag}_ncu_full.txt (ncu --set full, largest captured launch)"}
    json.dump(digest, open(path, "w"), indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 4:          # ... <windows_per_launch>: also refresh profiles/ncu_traffic.json
        traffic_digest(sys.argv[2], sys.argv[3], float(sys.argv[4]))
    launches(sys.argv[1], sys.argv[3] + "_launches.txt")
    if sys.argv[2] != "-":
        full(sys.argv[2], sys.argv[3] + "_ncu_full.txt")
