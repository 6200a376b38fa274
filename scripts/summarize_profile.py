#!/usr/bin/env python
"""Condense gpurun_out ncu artefacts into small tracked files under profiles/.

  summarize_profile.py launches <launches.csv> <out.md> [title]
  summarize_profile.py full <prof.ncu-rep> <out.md> [title]
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def short_name(name):
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", name)
    s = (m.group(1) + (m.group(2) or "")) if m else name[:60]
    if "at::" in name or "templates::" in name:
        return "torch (synthetic input generation / host-side tensor ops, outside the timed region)"
    return "vpint:" + s


def launches(path, out, title):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    seq = []
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        k = short_name(row["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        seq.append((k, v))
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as o:
        o.write("# %s\n\nSource: `ncu --metrics gpu__time_duration.sum --clock-control none` launch list (%d launches, "
                "%.1f ms of kernel time; serialised and cold-cache, so compare SHARES).\n\n" % (title, len(seq), tot / 1e3))
        o.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write("| `%s` | %d | %.1f | %.2f | %.1f%% |\n" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
        st = [v for k, v in seq if "jacobi" in k]
        if st:
            n = len(st)
            o.write("\nStencil launch duration along the run (launch index: us): "
                    + ", ".join("%d: %.1f" % (i, st[i]) for i in range(0, n, max(1, n // 24))) + "\n")


def full(path, out, title):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as o:
        o.write("# %s\n\nSource: `ncu --set full --clock-control none --import-source on`, read with "
                "`ncu -i ... --page raw --csv`.\n" % title)
        for r in rows[2:]:
            o.write("\n## `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % r[idx["Kernel Name"]])
            for k in KEEP:
                if k in idx:
                    o.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
            try:
                rd = float(r[idx["dram__bytes_read.sum"]]); wr = float(r[idx["dram__bytes_write.sum"]])
                ur, uw = units[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_write.sum"]]
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tb = rd * scale[ur] + wr * scale[uw]
                t = float(r[idx["gpu__time_duration.sum"]]); ut = units[idx["gpu__time_duration.sum"]]
                ts = t * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[ut]
                o.write("\nDRAM traffic %.1f MB per launch, %.0f GB/s over the launch.\n" % (tb / 1e6, tb / ts / 1e9))
            except Exception as e:  # pragma: no cover
                o.write("\n(traffic not computed: %s)\n" % e)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, dst, title)
