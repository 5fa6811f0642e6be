#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into small tracked summaries under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_x_launches.md
  python scripts/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r01_x_full.md
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ,
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`), source: %s\n\n" % src)
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | avg us | total us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in agg.items():
            f.write("| `%s` | %d | %.1f | %.1f | %.3f |\n" % (k[:90], len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, sum(v) / tot))
    print(open(dst).read())


def traffic(src, dst, key):
    """profiles/traffic.json: DRAM bytes (read + write) per launch of the captured kernel"""
    import json
    import os

    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = []
    for r in rows[2:]:
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(r[i].replace(",", "")) * scale[units[i]]
        vals.append(tot)
    t = json.load(open(dst)) if os.path.exists(dst) else {}
    t[key] = sum(vals) / len(vals)
    t[key + "_source"] = os.path.basename(src) + " (ncu --set full, cold L2: every launch re-reads its walkers from HBM)"
    json.dump(t, open(dst, "w"), indent=1)
    print(t)


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write("# ncu --set full summary, source: %s\n\n" % src)
        for r in rows[2:]:
            f.write("## `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % r[ki][:100])
            for m in FULL_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write("| %s | %s | %s |\n" % (m, r[i], units[i]))
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
